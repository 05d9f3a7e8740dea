#!/bin/bash
# patch kernels: CUDA's volatile-asm intrinsics (__dp2a_*, __funnelshift_rc) or plain-asm equivalents
cd /root/repo
for v in 0 1; do
  rm -f ffvvc_b200/csrc/build/inter_patch.o
  make -s -C ffvvc_b200/csrc EXTRA="-DINTER_PLAIN_ASM=$v" > /dev/null 2>&1
  echo -n "INTER_PLAIN_ASM=$v "
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_launch'],3) for k,v in d['roofline']['stages'].items() if k in ('inter',)}, round(d['value']), d['parity']['equal'])"
done
rm -f ffvvc_b200/csrc/build/inter_patch.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
