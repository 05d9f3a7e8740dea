#!/bin/bash
# Patch kernels: staged windows (PATCH_STAGED=1) against register staging (0); resident CTAs of the luma kernels
cd /root/repo
first=1
for cfg in ${PATCH_SWEEP:-"-DPATCH_STAGED=1" "-DPATCH_STAGED=0" "-DPATCH_MB_LB=3" "-DPATCH_MB_LB=3 -DPATCH_MB_LU=3"}; do
  rm -f ffvvc_b200/csrc/build/inter_patch.o
  if ! make -s -C ffvvc_b200/csrc EXTRA="$cfg" > /tmp/mk.log 2>&1; then echo "build failed for $cfg"; tail -3 /tmp/mk.log; continue; fi
  echo -n "$cfg "
  if [ $first = 1 ]; then echo -n "tests: $(timeout 600 python -m pytest tests/test_gpu_inter.py tests/test_gpu_recon.py -m gpu -x -q 2>&1 | tail -1) "; first=0; fi
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('inter',)}, round(d['value']), d['parity']['equal'])"
done
rm -f ffvvc_b200/csrc/build/inter_patch.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
