#!/bin/bash
# residual warp kernel: mid-stage row pitch (shared memory per warp) x resident CTAs per SM
cd /root/repo
for v in "40 7" "32 7" "32 8" "32 9"; do
  set -- $v
  rm -f ffvvc_b200/csrc/build/itx_warp.o
  make -s -C ffvvc_b200/csrc EXTRA="-DITX_MID_PITCH=$1 -DITX_WARP_MB=$2" > /dev/null 2>&1
  echo -n "ITX_MID_PITCH=$1 ITX_WARP_MB=$2 $(cuobjdump -res-usage ffvvc_b200/csrc/build/itx_warp.o 2>/dev/null | grep -A1 'itx_warp_kernelILi2' | grep -o 'REG:[0-9]*\|STACK:[0-9]*\|SHARED:[0-9]*' | tr '\n' ' ')"
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('residual',)}, round(d['value']), d['parity']['equal'])"
done
rm -f ffvvc_b200/csrc/build/itx_warp.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
