#!/bin/bash
# Round-2 end-of-session evidence on the GPU box: GPU test suite, default bench line, reference arm, launch list of the
# bench command, one `ncu --set full` capture of a reconstruction pass at 16 pictures per launch, config 1-4 lines.
cd /root/repo
tag=${1:-r02_final}
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/tests_$tag.log; cat gpurun_out/tests_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
python tools/bench_configs.py > gpurun_out/bench_configs_$tag.jsonl 2> gpurun_out/bench_configs_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --quick --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
python tools/profile_recon.py 16 1 > gpurun_out/profile_recon_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"inter_|itx_|lmcs_|deblock_|sao_|alf_" -c 40 -o gpurun_out/recon_$tag -f python tools/profile_recon.py 16 1 > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/*$tag*
cut -c1-300 gpurun_out/bench_$tag.json
