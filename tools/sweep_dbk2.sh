#!/bin/bash
# deblock thread-per-segment kernel: resident CTAs per SM the kernel is compiled for (bench stage times, ms per launch)
cd /root/repo
for n in ${DBK_SWEEP:-0 4 5 6 7 8}; do
  rm -f ffvvc_b200/csrc/build/deblock.o
  make -s -C ffvvc_b200/csrc EXTRA="-DDBK_MIN_CTAS=$n" > /dev/null 2>&1
  echo -n "DBK_MIN_CTAS=$n "
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k.startswith('deblock')}, round(d['value']), d['parity']['equal'])"
done
rm -f ffvvc_b200/csrc/build/deblock.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
