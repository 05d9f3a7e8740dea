#!/bin/bash
# SAO kernel: resident CTAs per SM it is compiled for (bench stage times, ms per launch)
cd /root/repo
for n in ${SAO_SWEEP:-3 4 5 6 8}; do
  rm -f ffvvc_b200/csrc/build/sao.o
  make -s -C ffvvc_b200/csrc EXTRA="-DSAO_MIN_CTAS=$n" > /dev/null 2>&1
  echo -n "SAO_MIN_CTAS=$n $(cuobjdump -res-usage ffvvc_b200/csrc/build/sao.o 2>/dev/null | grep -o 'REG:[0-9]*\|STACK:[0-9]*' | head -2 | tr '\n' ' ')"
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('sao',)}, round(d['value']), d['parity']['equal'])"
done
rm -f ffvvc_b200/csrc/build/sao.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
