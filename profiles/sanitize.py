#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python profiles/sanitize.py
Host buffers only (no torch): every call stages through the library's own device allocations."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ministark_b200 as ms
from ministark_b200 import expr as E

ctx = ms.Context(0)
rng = np.random.default_rng(0)
rand = lambda *shape: rng.integers(0, ms.P, size=shape, dtype=np.uint64)
for log_n in (3, 9, 12, 13, 16):                       # single-tile, two-pass and three-digit NTT plans, Fp and Fq3
    for field in (ms.FP, ms.FQ3):
        a = rand(3, field << log_n)
        b = a.copy()
        ctx.ntt_batch(b, field, log_n, 3, offset=ms.GENERATOR)
        ctx.ntt_batch(b, field, log_n, 3, inverse=True, offset=ms.GENERATOR)
        assert np.array_equal(a, b), (log_n, field)
        out = np.empty((3, field << (log_n + 2)), dtype=np.uint64)
        ctx.lde_batch(a, out, field, log_n, 2, 3, offset=ms.GENERATOR, bitrev=True)
n = 1 << 12
cols = rand(5, n)
root = ctx.merkle_commit(cols, ms.FP, n, 5)
leaves, nodes = np.empty((n, 32), dtype=np.uint8), np.empty((n, 32), dtype=np.uint8)
ctx.merkle_commit(cols, ms.FP, n, 5, leaves=leaves, nodes=nodes)
ctx.merkle_prove(leaves, nodes, n, [1, 5, 4, 4000])
ctx.merkle_commit_rows(rand(n // 8, 24), 24, n // 8)
ctx.gather_rows(cols, ms.FP, n, 5, [0, 7, n - 1])
for field in (ms.FP, ms.FQ3):
    for log_ff in (1, 3, 4):
        ev = rand(field << 12)
        out = np.empty(field << (12 - log_ff), dtype=np.uint64)
        ctx.fri_fold(ev, out, field, 12, log_ff, rand(3))
    for m in (1, 2049, 5000):
        out = np.empty(m * field, dtype=np.uint64)
        ctx.scan_affine(out, field, m, rand(field), a=rand(m * field), a_field=field, b=rand(m), b_field=ms.FP, inclusive=True)
x = E.X()
expr = (E.Trace(0, 1) - E.Trace(1, 0) * E.Trace(2, 0)) * (x - E.Constant(5)) / (x ** n - E.Constant(1)) + E.Trace(3, 0) ** 3
prog = E.compile_program(expr, 5, log_ce=12)
out = np.empty(n, dtype=np.uint64)
ctx.eval_constraints(prog, out, 12, base_cols=cols, nbase=5, fq_field=ms.FP, offset=ms.GENERATOR, trace_bitrev=True)
ctx.poly_eval(cols, ms.FP, n, 5, rand(2, 3))
acc = np.empty(n, dtype=np.uint64)
ctx.sum_columns(cols, acc, ms.FP, n, 5)
ctx.pow_grind(bytes(range(32)), 10)
print("sanitize.py: all kernel families ran; launches:", ctx.launches)
