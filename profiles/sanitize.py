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
# the persistent TMA pipeline (csrc/ntt_tma.cu) needs >= 1024 tiles of work: 2^16 points = digits [8, 8]
a_tma, w = rand(16, 1 << 16), rand(64, 1 << 16)
for tma in (1, 0):
    ctx.set_option("ntt_tma", tma)
    a = a_tma
    out_t = np.empty((16, 1 << 18), dtype=np.uint64)
    ctx.lde_batch(a, out_t, ms.FP, 16, 2, 16, offset=ms.GENERATOR, bitrev=True)          # strided (pre-scale) + contiguous pass
    if tma:
        want_lde = out_t.copy()
    else:
        assert np.array_equal(want_lde, out_t)
    v = w.copy()
    ctx.ntt_batch(v, ms.FP, 16, 64, offset=ms.GENERATOR)                                # strided pass, natural digit
    ctx.ntt_batch(v, ms.FP, 16, 64, inverse=True, offset=ms.GENERATOR)
    assert np.array_equal(v, w)
ctx.set_option("ntt_tma", 1)
try:        # scatter pass through per-block tensor maps (device buffers: torch is only needed for this leg)
    import torch
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        c2 = ms.Context(0, stream=st.cuda_stream)
        co = torch.from_numpy(rand(32, 1 << 16).view(np.int64)).cuda()
        work = torch.zeros((32, 1 << 19), dtype=torch.int64, device="cuda")
        slabs = [torch.zeros((32, 1 << 18), dtype=torch.int64, device="cuda") for _ in range(2)]
        blocks = [slabs[q // 4].data_ptr() + (q % 4) * (1 << 16) * 8 for q in range(8)]
        c2.lde_batch_scatter(co, work, ms.FP, 16, 3, 32, blocks, 1 << 18, [work.data_ptr()] + [0] * 7, 1 << 19)
        c2.sync()
        ref = torch.empty((32, 1 << 19), dtype=torch.int64, device="cuda")
        c2.lde_batch(co, ref, ms.FP, 16, 3, 32)
        c2.sync()
        assert torch.equal(torch.cat([slabs[0], slabs[1]], dim=1), ref)
except ImportError:
    pass
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
# periodic column + the SUB opcode through both evaluators
pexpr = E.Trace(0, 0) * E.Periodic([3, 5, 7, 11], 8) - E.X()
pprog = E.compile_program(pexpr, 5, log_ce=12, num_cols=5)
tabs = E.periodic_tables(ctx, pprog, 12, 1)
try:
    import torch
    dcols = torch.from_numpy(cols.view(np.int64)).cuda()
    dout = torch.empty(n, dtype=torch.int64, device="cuda")
    for env in (None, "1"):
        if env:
            os.environ["MS_EVAL_NO_JIT"] = env
        ctx.eval_constraints_ptrs(pprog, dout, 12, [dcols[i] for i in range(5)] + [p for p, _ in tabs], [False] * 5 + [q for _, q in tabs],
                                  fq_field=ms.FP)
        ctx.sync()
    os.environ.pop("MS_EVAL_NO_JIT", None)
except ImportError:
    pass
for p_, _ in tabs:
    ctx.free(p_)
ctx.poly_eval(cols, ms.FP, n, 5, rand(2, 3))
acc = np.empty(n, dtype=np.uint64)
ctx.sum_columns(cols, acc, ms.FP, n, 5)
ctx.pow_grind(bytes(range(32)), 10)
print("sanitize.py: all kernel families ran; launches:", ctx.launches)
