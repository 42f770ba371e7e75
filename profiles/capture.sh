#!/bin/bash
# Profiling recipe (B200_PROFILING.md) — run under gpurun from the repo root; outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
# 1. every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/launches_bench.log 2>&1
# 2. full-set captures of the top kernels on a smaller instance of the same pipeline (2^22 rows)
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 3 -c 3 -o gpurun_out/prof_ntt_lde_r01 \
    python bench.py --steps 1 --warmup 0 --no-cpu --log-n 22 > gpurun_out/prof_ntt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hash_rows_kernel|eval_kernel|merkle_level_kernel" -c 3 \
    -o gpurun_out/prof_hash_eval_r01 python bench.py --steps 1 --warmup 0 --no-cpu --log-n 22 > gpurun_out/prof_hash.log 2>&1
ls -la gpurun_out
