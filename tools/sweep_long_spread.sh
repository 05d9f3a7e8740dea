#!/bin/bash
# Long launches: small inter kernels (chroma classes, PROF) on a side stream or not (bench stage times, ms per launch)
cd /root/repo
for v in ${SPREAD_SWEEP:-0 1}; do
  rm -f ffvvc_b200/csrc/build/inter.o
  make -s -C ffvvc_b200/csrc EXTRA="-DINTER_LONG_SPREAD=$v" > /dev/null 2>&1
  echo -n "INTER_LONG_SPREAD=$v "
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('inter','residual')}, round(d['value']))"
done
rm -f ffvvc_b200/csrc/build/inter.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
