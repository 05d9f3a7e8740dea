#!/bin/bash
# DMVR / BDOF warp kernels: window rows a lane requests before it stores the first one
cd /root/repo
for v in 4 8 12 23; do
  rm -f ffvvc_b200/csrc/build/inter_warp.o
  make -s -C ffvvc_b200/csrc EXTRA="-DINTER_STAGE_UNROLL=$v" > /dev/null 2>&1
  echo -n "INTER_STAGE_UNROLL=$v $(cuobjdump -res-usage ffvvc_b200/csrc/build/inter_warp.o 2>/dev/null | grep -A1 'inter_warp_kernelILi0ELi0ELi[12]' | grep -o 'REG:[0-9]*' | tr '\n' ' ')"
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('inter',)}, round(d['value']), d['parity']['equal'])"
done
rm -f ffvvc_b200/csrc/build/inter_warp.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
