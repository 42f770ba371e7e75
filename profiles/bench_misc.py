#!/usr/bin/env python
"""Side measurements on one B200 (JSON lines):
  config1  — BASELINE config 1: 2^20-point forward + inverse NTT, Fp, ONE column (resident, CUDA events)
  batchntt — 32 columns x 2^20..2^24, forward NTT, resident
  lde_fp / lde_fq3 — the same number of words as Fp columns and as Fq3 columns (LDE x16, 2^20 rows)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ministark_b200 as ms

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = ms.Context(0, stream=stream.cuda_stream)


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


n = 1 << 20
col = torch.empty(n, dtype=torch.int64, device="cuda")
ctx.fill_random(col, n, 1)
for off, nm in ((ms.ONE, "subgroup"), (ms.GENERATOR, "coset")):
    t = timed(lambda: (ctx.ntt_batch(col, ms.FP, 20, offset=off), ctx.ntt_batch(col, ms.FP, 20, inverse=True, offset=off)))
    print(json.dumps({"bench": "config1", "domain": nm, "fwd_plus_inv_ms": t, "field_ops_per_s": 2 * 1.5 * n * 20 / t * 1e3,
                      "algorithmic_GBps": 2 * 16 * n / t / 1e6}))
for log_n in (20, 22, 24):
    m = torch.empty((32, 1 << log_n), dtype=torch.int64, device="cuda")
    ctx.fill_random(m, 32 << log_n, 2)
    t = timed(lambda: ctx.ntt_batch(m, ms.FP, log_n, 32), reps=5)
    print(json.dumps({"bench": "batchntt", "log_n": log_n, "ncols": 32, "ms": t,
                      "field_ops_per_s": 32 * 1.5 * (1 << log_n) * log_n / t * 1e3, "algorithmic_GBps": 32 * 16 * (1 << log_n) / t / 1e6}))
    del m
log_n, log_b = 20, 4
for field, ncols, nm in ((ms.FP, 27, "lde_fp"), (ms.FQ3, 9, "lde_fq3")):
    src = torch.empty((ncols, field << log_n), dtype=torch.int64, device="cuda")
    dst = torch.empty((ncols, field << (log_n + log_b)), dtype=torch.int64, device="cuda")
    ctx.fill_random(src, ncols * field << log_n, 3)
    t = timed(lambda: ctx.lde_batch(src, dst, field, log_n, log_b, ncols), reps=5)
    print(json.dumps({"bench": nm, "log_n": log_n, "blowup": 1 << log_b, "ncols": ncols, "words": ncols * field << log_n, "ms": t}))
