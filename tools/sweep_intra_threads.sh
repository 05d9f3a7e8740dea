#!/bin/bash
# dependency-driven intra kernel: threads per CTA (the pictures' critical path is a chain of steps, so a step's latency counts)
cd /root/repo
for n in ${INTRA_SWEEP:-128 256 512}; do
  rm -f ffvvc_b200/csrc/build/intra.o
  if ! make -s -C ffvvc_b200/csrc EXTRA="-DINTRA_THREADS=$n" > /tmp/mk.log 2>&1; then echo "build failed for $n"; tail -3 /tmp/mk.log; continue; fi
  echo "INTRA_THREADS=$n"
  timeout 300 python tools/bench_intra.py 1920 1080 8 2>&1 | grep workload | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  x8', round(d['launch_per_wave']['ms_per_ring'],2), {k:(round(v['ms_per_ring'],2), v['parity_equal']) for k,v in d['one_launch'].items()})"
  timeout 300 python tools/bench_intra.py 1920 1080 1 2>&1 | grep workload | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  x1', round(d['launch_per_wave']['ms_per_ring'],2), {k:(round(v['ms_per_ring'],2), v['parity_equal']) for k,v in d['one_launch'].items()})"
done
rm -f ffvvc_b200/csrc/build/intra.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
