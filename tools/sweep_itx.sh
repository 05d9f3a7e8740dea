#!/bin/bash
# Sweep the compiled occupancy (resident CTAs per SM) of the residual warp kernel on the GPU box; prints the
# residual / inter stage times of bench.py (ms per launch of 8 pictures) for each setting.
cd /root/repo
for mb in ${ITX_SWEEP:-1 8 10 12}; do
  rm -f ffvvc_b200/csrc/build/itx_warp.o
  make -s -C ffvvc_b200/csrc EXTRA="-DITX_WARP_MB=$mb" > /dev/null 2>&1
  echo -n "ITX_WARP_MB=$mb "
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items()}, round(d['value']))"
done
rm -f ffvvc_b200/csrc/build/itx_warp.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
