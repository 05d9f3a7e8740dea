#!/bin/bash
# End-of-session evidence run on the GPU box: full GPU test suite, the default bench line, the reference arm, the
# launch list of the bench command and one `ncu --set full` capture of a whole reconstruction pass.
# Usage (under gpurun): tools/final_profiles.sh <tag>
cd /root/repo
tag=${1:-final}
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py > gpurun_out/ncu_bench_$tag.log 2>&1
python tools/profile_recon.py 2 2 > gpurun_out/profile_recon_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"inter_|itx_|lmcs_|deblock_|sao_|alf_" -s 17 -c 16 -o gpurun_out/recon_$tag -f python tools/profile_recon.py 2 2 > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/*$tag*
cut -c1-400 gpurun_out/bench_$tag.json
