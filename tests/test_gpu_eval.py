"""GPU parity: fused constraint evaluation (csrc/eval.cu) vs the CPU oracle's eval_cpu restatement.
Cases follow the reference's own eval tests (src/eval_gpu.rs:917-1082: X-only, mixed Fp/Fq3,
inverse, trace offsets, constants; trace_len 2048, blowup 1 and 4) plus the synthetic config-3 AIR."""
import random

import numpy as np
import pytest

import ministark_b200 as ms
from ministark_b200 import expr as E
from ministark_b200 import synth_air

pytestmark = pytest.mark.gpu
P = ms.P


@pytest.fixture(scope="module")
def ctx():
    return ms.Context(0)


def _run(ctx, orc, expr, log_m, base, ext, fq, challenges=(), hints=(), lde_step=1, bitrev=False):
    from oracle import eval_oracle
    nbase = 0 if base is None else base.shape[0]
    next_ = 0 if ext is None else ext.shape[0]
    prog = E.compile_program(expr, nbase, challenges, hints, lde_step=lde_step, log_ce=log_m)
    out = np.empty((1 << log_m) * fq, dtype=np.uint64)
    b, e = base, ext
    if bitrev:  # hand the columns over in bit-reversed storage order
        b = None if base is None else np.stack([orc.bit_reverse(c, 1, log_m) for c in base])
        e = None if ext is None else np.stack([orc.bit_reverse(c, fq, log_m) for c in ext])
    ctx.eval_constraints(prog, out, log_m, base_cols=b, nbase=nbase, ext_cols=e, next_=next_, fq_field=fq,
                         offset=ms.GENERATOR, trace_bitrev=bitrev)
    want = eval_oracle.evaluate(expr.to_tuple(), log_m, orc.generator(), base, ext, fq_lanes=fq,
                                challenges=challenges, hints=hints, lde_step=lde_step)
    assert np.array_equal(out, want)
    return out


@pytest.mark.parametrize("bitrev", [False, True])
@pytest.mark.parametrize("blowup", [1, 4])
def test_reference_style_cases(ctx, orc, bitrev, blowup):
    log_m = 11 + (blowup.bit_length() - 1)
    m = 1 << log_m
    rng = random.Random(7)
    base = orc.rand_matrix(3, m, 1, seed=1)
    ext = orc.rand_matrix(2, m, 3, seed=2)
    chal = [tuple(rng.randrange(P) for _ in range(3)) for _ in range(2)]
    hint = [tuple(rng.randrange(P) for _ in range(3))]
    x = E.X()
    cases = [
        x,                                                        # X only
        x * x + 5,                                                # constants
        (x ** 3 - 1) / (x - 3),                                   # inverse of an Fp expression
        E.Trace(0, 0) * E.Trace(1, 1) - E.Trace(2, -1),           # trace offsets, Fp
        E.Trace(3, 0) * E.Trace(0, 1) + E.Trace(4, 2) * E.Trace(3, -2),   # mixed Fp / Fq3
        (E.Trace(3, 1) + E.Challenge(0)) / (E.Trace(4, 0) - E.Hint(0)),   # Fq3 inverse
        E.Constant((1, 2, 3)) * x ** 5 - E.Challenge(1) ** 3,
        -(E.Trace(1, 0) ** 7) + E.Constant(9),
        E.Constant(4) * E.Constant(5) + E.Challenge(0) / E.Challenge(1),   # folds to a constant
    ]
    for ex in cases:
        _run(ctx, orc, ex, log_m, base, ext, 3, chal, hint, lde_step=blowup, bitrev=bitrev)


def test_fq_equals_fp_air(ctx, orc):
    # AIRs with Fq = Fp (examples/fib): challenges are base-field elements, output is one word per point
    log_m = 12
    base = orc.rand_matrix(2, 1 << log_m, 1, seed=5)
    ex = (E.Trace(0, 1) - E.Trace(0, 0) * E.Trace(1, 0)) * (E.X() ** 2 * E.Challenge(0) + E.Challenge(1))
    _run(ctx, orc, ex, log_m, base, None, 1, challenges=[11, 12])


@pytest.mark.parametrize("log_n,ncols", [(11, 32), (13, 8)])
def test_synthetic_config3_air(ctx, orc, log_n, ncols):
    from oracle import synth_oracle
    log_b = 3
    trace = orc.rand_matrix(ncols, 1 << log_n, 1, seed=9)
    polys = orc.ntt(trace, 1, log_n, inverse=True)
    lde = orc.lde(polys, 1, log_n, log_b, orc.generator(), bitrev=True)
    ev = synth_air.GpuConstraintEval(ctx, log_n, log_b, ncols)
    out = np.empty(1 << log_n, dtype=np.uint64)
    ev.run(lde, out)
    want = synth_oracle.constraint_eval(orc, lde, log_n, log_b, ncols)
    assert np.array_equal(out, want)
    # the product's tree and the oracle's independently stated tree are the same expression
    assert synth_air.composition(ncols, log_n).to_tuple() == synth_oracle.composition_tree(ncols, log_n)


def test_program_validation(ctx):
    prog = E.compile_program(E.Trace(5, 0) + 1, 8, log_ce=4)
    out = np.empty(16, dtype=np.uint64)
    base = np.zeros((2, 16), dtype=np.uint64)
    with pytest.raises(ms.MsError):          # column 5 referenced, only 2 provided (panics in eval_cpu.rs:148)
        ctx.eval_constraints(prog, out, 4, base_cols=base, nbase=2)


@pytest.mark.parametrize("bitrev", [False, True])
@pytest.mark.parametrize("blowup", [1, 4])
def test_periodic_columns(ctx, orc, bitrev, blowup):
    """Periodic(coeffs, interval) leaves (src/constraints.rs:107-146, src/eval_cpu.rs:176-256): tables built on the device
    (expr.periodic_tables), evaluator (interpreter and run-time specialised kernel) vs the oracle's definition"""
    torch = pytest.importorskip("torch")
    from oracle import eval_oracle
    log_n = 9
    log_m = log_n + (blowup.bit_length() - 1)
    m = 1 << log_m
    base = orc.rand_matrix(2, m, 1, seed=21)
    ext = orc.rand_matrix(1, m, 3, seed=22)
    pa = E.Periodic([3, 5, 11, 2], 8)
    pb = E.Periodic([(1, 2, 3), (4, 5, 6)], 64)
    ex = (E.Trace(0, 1) - pa * E.Trace(1, 0)) * pb + E.Trace(2, 0) * pa * pa - E.X() * pb
    prog = E.compile_program(ex, 2, lde_step=blowup, log_ce=log_m, num_cols=3)
    cols_np = [base[0], base[1], ext[0]]
    if bitrev:
        cols_np = [orc.bit_reverse(base[0], 1, log_m), orc.bit_reverse(base[1], 1, log_m), orc.bit_reverse(ext[0], 3, log_m)]
    dev = [torch.from_numpy(c.view(np.int64).copy()).cuda() for c in cols_np]
    tabs = E.periodic_tables(ctx, prog, log_n, blowup)
    out = torch.empty(m * 3, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    want = eval_oracle.evaluate(ex.to_tuple(), log_m, orc.generator(), base, ext, fq_lanes=3, lde_step=blowup)
    import os
    for no_jit in (False, True):
        if no_jit:
            os.environ["MS_EVAL_NO_JIT"] = "1"
        try:
            out.zero_()
            torch.cuda.synchronize()
            ctx.eval_constraints_ptrs(prog, out, log_m, dev + [p for p, _ in tabs], [False, False, True] + [q for _, q in tabs],
                                      fq_field=3, trace_bitrev=bitrev)
            ctx.sync()
        finally:
            os.environ.pop("MS_EVAL_NO_JIT", None)
        assert np.array_equal(out.cpu().numpy().view(np.uint64), want)
    for p, _ in tabs:
        ctx.free(p)


def test_sub_opcode_and_register_validation(ctx):
    """OP_SUB (declared in the instruction set, emitted by the C++ host compiler) in both evaluators, and the
    defined-before-use check of source registers"""
    torch = pytest.importorskip("torch")
    import os
    log_m = 6
    m = 1 << log_m
    c5 = 5 * 2**64 % P
    consts = np.array([[c5, 0, 0]], dtype=np.uint64)
    code = np.array([[E.OP_X, 0, 0, 0], [E.OP_CONST, 1, 0, 0], [E.OP_SUB, 2, 1, 0], [E.OP_STORE, 0, 2, 0]], dtype=np.uint32)
    prog = E.Program(code, consts, 3, False)
    g = pow(pow(7, (P - 1) >> 32, P), 1 << (32 - log_m), P)
    want = np.array([((5 - 7 * pow(g, i, P)) % P) * 2**64 % P for i in range(m)], dtype=np.uint64)
    out = torch.empty(m, dtype=torch.int64, device="cuda")
    for no_jit in (False, True):
        if no_jit:
            os.environ["MS_EVAL_NO_JIT"] = "1"
        try:
            out.zero_()
            torch.cuda.synchronize()
            ctx.eval_constraints_ptrs(prog, out, log_m, [], [], fq_field=1)
            ctx.sync()
        finally:
            os.environ.pop("MS_EVAL_NO_JIT", None)
        assert np.array_equal(out.cpu().numpy().view(np.uint64), want)
    bad = E.Program(np.array([[E.OP_X, 0, 0, 0], [E.OP_ADD, 2, 0, 7], [E.OP_STORE, 0, 2, 0]], dtype=np.uint32), consts, 3, False)
    with pytest.raises(ms.MsError):          # register 7 is read but never written
        ctx.eval_constraints_ptrs(bad, out, log_m, [], [], fq_field=1)
