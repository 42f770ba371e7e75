#!/bin/bash
# What to run first on the next B200 slot: everything that changed AFTER the round-2 GPU budget was spent and is therefore
# verified on the CPU only (DESIGN.md §8.1 "Fewer field operations per point", the host-path trims, the device-built fib
# trace).  From the repo root:
#     gpurun --timeout 1500 -- 'bash profiles/remeasure_after_r02.sh r03a'
# Outputs in gpurun_out/.  Expectations to check the results against are in profiles/jit_static_r02.md.
TAG=${1:-r03a}
set -x
mkdir -p gpurun_out
# 1. parity first: the whole GPU suite (the prover tests compare bytes with the CPU restatement; the three tests of the
#    device-built fib trace run last by name)
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_${TAG}.log
# 2. the programs the compiler rewrites touched: brainfuck hello_world (latency: host trims) and the 2^20-row trace
#    (throughput: composition 54 -> expected ~41 ms, DEEP 45 -> expected ~19 ms by multiplication count)
python profiles/bench_brainfuck.py > gpurun_out/bench_brainfuck_${TAG}.json 2> gpurun_out/bench_brainfuck_${TAG}.log
python profiles/bench_brainfuck.py --burner 40 40 60 > gpurun_out/bench_brainfuck_burner_${TAG}.json 2>> gpurun_out/bench_brainfuck_${TAG}.log
python profiles/bench_prover.py > gpurun_out/bench_prover_${TAG}.jsonl 2> gpurun_out/bench_prover_${TAG}.log
# 3. per-launch device times of one 2^20-row brainfuck proof (the two ms_eval_jit launches over 2^24 points are the
#    composition and the DEEP kernels) and a full-set capture of them
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_bf_${TAG}.csv python profiles/bench_brainfuck.py --burner 40 40 60 > gpurun_out/launches_bf_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ms_eval_jit -c 40 -o gpurun_out/prof_eval_bf_${TAG} \
    python profiles/bench_brainfuck.py --burner 16 16 14 > gpurun_out/prof_eval_bf_${TAG}.log 2>&1
# 4. the headline line, unchanged code path (its evaluator program is byte-identical to the verified one): a sanity line
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.log
ls -la gpurun_out | tail -12
