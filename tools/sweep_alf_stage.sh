#!/bin/bash
# ALF and deblock: tile staging by cp.async (ALF_STAGE_ASYNC / DBK_STAGE_ASYNC = 1) against load + store (0)
cd /root/repo
first=1
for cfg in ${ALF_SWEEP:-"-DALF_STAGE_ASYNC=1 -DDBK_STAGE_ASYNC=1" "-DALF_STAGE_ASYNC=0 -DDBK_STAGE_ASYNC=0"}; do
  rm -f ffvvc_b200/csrc/build/alf.o ffvvc_b200/csrc/build/deblock.o
  if ! make -s -C ffvvc_b200/csrc EXTRA="$cfg" > /tmp/mk.log 2>&1; then echo "build failed for $cfg"; tail -3 /tmp/mk.log; continue; fi
  echo -n "$cfg "
  if [ $first = 1 ]; then echo -n "tests: $(timeout 600 python -m pytest tests/test_gpu_alf.py tests/test_gpu_inloop.py tests/test_gpu_lf_sao.py -m gpu -x -q 2>&1 | tail -1) "; first=0; fi
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('alf','deblock_v','deblock_h')}, round(d['value']), d['parity']['equal'])"
done
rm -f ffvvc_b200/csrc/build/alf.o ffvvc_b200/csrc/build/deblock.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
