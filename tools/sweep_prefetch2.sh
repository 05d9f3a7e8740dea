#!/bin/bash
# DMVR / BDOF warp kernels: next record's reference windows prefetched into L2 or not
cd /root/repo
for v in 0 1; do
  rm -f ffvvc_b200/csrc/build/inter_warp.o
  make -s -C ffvvc_b200/csrc EXTRA="-DINTER_WARP_PREFETCH=$v" > /dev/null 2>&1
  echo -n "INTER_WARP_PREFETCH=$v "
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('inter',)}, round(d['value']), d['parity']['equal'])"
done
rm -f ffvvc_b200/csrc/build/inter_warp.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
