#!/bin/bash
# Round-2 evidence run on the GPU box: launch list of the bench command and one `ncu --set full` capture of a whole
# reconstruction pass at the bench's 16 pictures per launch (so that the per-kernel DRAM traffic is measured on a working
# set far above the 126 MB L2).  Usage (under gpurun): tools/r02_profiles.sh <tag>
cd /root/repo
tag=${1:-r02}
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_$tag.log 2>&1
python tools/profile_recon.py 16 1 > gpurun_out/profile_recon_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"inter_|itx_|lmcs_|deblock_|sao_|alf_" -c 40 -o gpurun_out/recon_$tag -f python tools/profile_recon.py 16 1 > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/*$tag*
cut -c1-300 gpurun_out/bench_$tag.json
