#!/bin/bash
# Round-2 profiling recipe (B200_PROFILING.md) — run under gpurun from the repo root:
#     gpurun -- 'bash profiles/capture_r02.sh r02b'
# outputs land in gpurun_out/; summaries are then written to profiles/ by `python profiles/summarize.py r02b`.
TAG=${1:-r02}
set -x
mkdir -p gpurun_out
# 1. every launch of the full-size step (2^24 x 32) with device time and DRAM bytes
#    (cold-cache, serialised: compare SHARES with the CUDA-event phase times, not absolutes)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --no-verify --no-prover \
    > gpurun_out/launches_bench_${TAG}.log 2>&1
# 2. full-set capture of the NTT pipeline kernels at the full transform size (2^24 points per coset, 4 columns x 8 cosets)
ncu --set full --clock-control none --import-source on -k regex:ntt_tma -s 3 -c 3 -o gpurun_out/prof_ntt_lde_${TAG} \
    python profiles/exp_ntt_r02.py prof > gpurun_out/prof_ntt_${TAG}.log 2>&1
# 3. full-set capture of the hash / evaluator kernels on a 2^22-row instance of the step
ncu --set full --clock-control none --import-source on -k regex:"hash_rows_kernel|ms_eval_jit|merkle_level_kernel" -c 3 \
    -o gpurun_out/prof_hash_eval_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-extra --no-verify --no-prover --log-n 22 \
    > gpurun_out/prof_hash_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
