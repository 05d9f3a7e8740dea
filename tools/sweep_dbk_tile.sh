#!/bin/bash
# deblock tile shapes (V pass width x height, H pass height, width) (the apron and the per-CTA set-up are amortised over more samples)
cd /root/repo
IFS=";" read -ra cfgs <<< "${DBK_SWEEP:-128 32 32 128;256 32 32 128;128 64 32 128;128 32 64 128;128 32 32 256;128 32 64 256}"
for cfg in "${cfgs[@]}"; do
  IFS=" " read -r a b c d <<< "$cfg"; set -- $a $b $c $d
  rm -f ffvvc_b200/csrc/build/deblock.o
  if ! make -s -C ffvvc_b200/csrc EXTRA="-DDBK_TW_V=$1 -DDBK_TH_V=$2 -DDBK_TH_H=$3 -DDBK_TW_H=$4" > /tmp/mk.log 2>&1; then echo "build failed for $cfg"; tail -3 /tmp/mk.log; continue; fi
  echo -n "V ${1}x$2 H ${4}x$3 "
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k.startswith('deblock')}, round(d['value']), d['parity']['equal'])"
done
rm -f ffvvc_b200/csrc/build/deblock.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
