#!/bin/bash
# SAO: rows per lane (the halo rows and the per-warp set-up are amortised over more rows)
cd /root/repo
for n in ${SAO_SWEEP:-4 8 16}; do
  rm -f ffvvc_b200/csrc/build/sao.o
  if ! make -s -C ffvvc_b200/csrc EXTRA="-DSAO_RPT=$n" > /tmp/mk.log 2>&1; then echo "build failed for $n"; tail -3 /tmp/mk.log; continue; fi
  echo -n "SAO_RPT=$n tests: $(timeout 300 python -m pytest tests/test_gpu_lf_sao.py tests/test_gpu_inloop.py -m gpu -x -q 2>&1 | tail -1) "
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('sao',)}, round(d['value']), d['parity']['equal'])"
done
rm -f ffvvc_b200/csrc/build/sao.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
