#!/bin/bash
# Sweep the resident CTAs per SM of the inter warp kernels (DMVR / BDOF / PROF records) on the GPU box
cd /root/repo
for n in ${CTAS_SWEEP:-3 4 5 6 7}; do
  rm -f ffvvc_b200/csrc/build/inter_warp.o
  make -s -C ffvvc_b200/csrc EXTRA="-DINTER_WARP_CTAS=$n" > /dev/null 2>&1
  echo -n "INTER_WARP_CTAS=$n "
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('inter','residual')}, round(d['value']))"
done
rm -f ffvvc_b200/csrc/build/inter_warp.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
