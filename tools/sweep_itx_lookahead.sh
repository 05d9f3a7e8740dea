#!/bin/bash
# Residual warp kernel: work items drawn two ahead, the next item's first block requested during this item's last block
cd /root/repo
first=1
for cfg in ${ITX_SWEEP:-"-DITX_LOOKAHEAD=1" "-DITX_LOOKAHEAD=0"}; do
  rm -f ffvvc_b200/csrc/build/itx_warp.o
  if ! make -s -C ffvvc_b200/csrc EXTRA="$cfg" > /tmp/mk.log 2>&1; then echo "build failed for $cfg"; tail -3 /tmp/mk.log; continue; fi
  echo -n "$cfg "
  if [ $first = 1 ]; then echo -n "tests: $(timeout 600 python -m pytest tests/test_gpu_itx.py tests/test_gpu_recon.py -m gpu -x -q 2>&1 | tail -1) "; first=0; fi
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('residual',)}, round(d['value']), d['parity']['equal'])"
done
rm -f ffvvc_b200/csrc/build/itx_warp.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
