#!/bin/bash
# Warp kernels of the inter stage: window rows by cp.async (INTER_STAGE_ASYNC=1) against load + store (0)
cd /root/repo
first=1
for cfg in ${STAGE_SWEEP:-"-DINTER_STAGE_ASYNC=1" "-DINTER_STAGE_ASYNC=0"}; do
  rm -f ffvvc_b200/csrc/build/inter_warp.o
  if ! make -s -C ffvvc_b200/csrc EXTRA="$cfg" > /tmp/mk.log 2>&1; then echo "build failed for $cfg"; tail -3 /tmp/mk.log; continue; fi
  echo -n "$cfg "
  if [ $first = 1 ]; then echo -n "tests: $(timeout 600 python -m pytest tests/test_gpu_inter.py tests/test_gpu_recon.py -m gpu -x -q 2>&1 | tail -1) "; first=0; fi
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('inter',)}, round(d['value']), d['parity']['equal'])"
done
rm -f ffvvc_b200/csrc/build/inter_warp.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
