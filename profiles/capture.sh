#!/bin/bash
# Profiling recipe (B200_PROFILING.md) — run under gpurun from the repo root:
#     gpurun -- 'bash profiles/capture.sh r01b'
# outputs land in gpurun_out/; summaries are then written to profiles/ by profiles/summarize.py.
TAG=${1:-r01}
set -x
mkdir -p gpurun_out
# 1. every launch of the full-size step (2^24 x 32) with device time and DRAM bytes
#    (cold-cache, serialised: compare SHARES with the CUDA-event phase times, not absolutes)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/launches_bench_${TAG}.log 2>&1
# 2. full-set captures of the top kernels on a smaller instance of the same pipeline (2^22 rows)
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 3 -c 3 -o gpurun_out/prof_ntt_lde_${TAG} \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --log-n 22 > gpurun_out/prof_ntt_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hash_rows_kernel|eval_kernel|merkle_level_kernel" -c 3 \
    -o gpurun_out/prof_hash_eval_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --log-n 22 > gpurun_out/prof_hash_${TAG}.log 2>&1
ls -la gpurun_out
