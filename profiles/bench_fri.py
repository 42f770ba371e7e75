#!/usr/bin/env python
"""BASELINE config 4: FRI fold sweep over Fq3 codewords 2^20..2^26 (ff = 8 and 16), one B200.
Prints one JSON line per (log_n, ff): device time of ms_fri_fold and of the in-place row commitment,
algorithmic GB/s (s*M + s*M/ff) against the measured HBM peak."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ministark_b200 as ms

peak = 6564.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = ms.Context(0, stream=stream.cuda_stream)
alpha = np.array([ms.to_mont(3), ms.to_mont(5), ms.to_mont(7)], dtype=np.uint64)
for log_n in range(20, 27, 2):
    n = 1 << log_n
    ev = torch.empty(3 * n, dtype=torch.int64, device="cuda")
    ctx.fill_random(ev, 3 * n, log_n)
    for log_ff in (3, 4):
        out = torch.empty(3 * n >> log_ff, dtype=torch.int64, device="cuda")
        leaves = torch.empty((n >> log_ff, 4), dtype=torch.int64, device="cuda")
        nodes = torch.empty((n >> log_ff, 4), dtype=torch.int64, device="cuda")
        for _ in range(3):
            ctx.fri_fold(ev, out, ms.FQ3, log_n, log_ff, alpha)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps = 10
        e[0].record()
        for _ in range(reps):
            ctx.fri_fold(ev, out, ms.FQ3, log_n, log_ff, alpha)
        e[1].record()
        for _ in range(reps):
            ctx.merkle_commit_rows(ev, 3 << log_ff, n >> log_ff, leaves=leaves, nodes=nodes)
        e[2].record()
        torch.cuda.synchronize()
        t_fold, t_commit = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
        alg = 24 * n + 24 * (n >> log_ff)
        print(json.dumps({"log_n": log_n, "ff": 1 << log_ff, "fold_ms": t_fold, "fold_GBps": alg / t_fold / 1e6,
                          "fold_frac_of_hbm_peak": alg / t_fold / 1e6 / peak, "commit_ms": t_commit}))
