#!/bin/bash
# DMVR / BDOF records as one kernel or as two (search, then motion compensation + BDOF + chroma); resident CTAs of the search kernel
cd /root/repo
for v in "0 7" "1 7" "1 8" "1 9"; do
  set -- $v
  rm -f ffvvc_b200/csrc/build/inter_warp.o
  make -s -C ffvvc_b200/csrc EXTRA="-DINTER_WARP_SPLIT=$1 -DINTER_WARP_CTAS_SEARCH=$2" > /dev/null 2>&1
  echo -n "INTER_WARP_SPLIT=$1 CTAS_SEARCH=$2 "
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('inter',)}, round(d['value']), d['parity']['equal'], d['gpu_launches'])"
done
rm -f ffvvc_b200/csrc/build/inter_warp.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
