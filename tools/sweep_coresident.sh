#!/bin/bash
# Experiment: the DMVR / BDOF warp kernel compiled for fewer resident CTAs per SM so that the latency-bound patch kernels,
# issued on parallel streams, can share the SMs with it (long launches; bench stage times, ms per launch)
cd /root/repo
for n in 3 4 5 7; do
  rm -f ffvvc_b200/csrc/build/inter_warp.o ffvvc_b200/csrc/build/inter.o
  make -s -C ffvvc_b200/csrc EXTRA="-DINTER_WARP_CTAS=$n -DINTER_LONG_SPREAD=3" > /dev/null 2>&1
  echo -n "INTER_WARP_CTAS=$n INTER_LONG_SPREAD=3 "
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('inter','residual')}, round(d['value']))"
done
rm -f ffvvc_b200/csrc/build/inter_warp.o ffvvc_b200/csrc/build/inter.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
