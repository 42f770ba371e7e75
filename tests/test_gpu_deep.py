"""GPU parity: DEEP composition (src/composer.rs) — OOD evaluations by parallel reduction and the DEEP
polynomial evaluated pointwise over the LDE — against the oracle's restatement of the reference's
coefficient-form path (Horner, synthetic division, column sum, degree adjustment, LDE)."""
import random

import numpy as np
import pytest

import ministark_b200 as ms
from ministark_b200 import deep
from ministark_b200 import expr as E

pytestmark = pytest.mark.gpu
P = ms.P


@pytest.fixture(scope="module")
def ctx():
    return ms.Context(0)


def _mont3(v):
    return np.array([ms.to_mont(c) for c in v], dtype=np.uint64)


@pytest.mark.parametrize("field,log_n", [(1, 3), (1, 10), (1, 14), (3, 9), (1, 15), (3, 15)])
def test_poly_eval_vs_horner(ctx, orc, field, log_n):
    n, ncols = 1 << log_n, 3
    rng = random.Random(log_n)
    coeffs = orc.rand_matrix(ncols, n, field, seed=log_n)
    pts = [tuple(rng.randrange(P) for _ in range(3)) for _ in range(3)] + [(5, 0, 0)]
    got = ctx.poly_eval(coeffs, field, n, ncols, np.stack([_mont3(p) for p in pts]))
    for c in range(ncols):
        for k, p in enumerate(pts):
            assert np.array_equal(got[c, k], orc.horner(coeffs[c], field, _mont3(p))), (c, k)


def test_poly_eval_multi_level(ctx, orc):
    # n > 256*64*... two reduction levels (2^24-row traces take two as well)
    n = 1 << 18
    coeffs = orc.rand_matrix(2, n, 1, seed=1)
    p = (123456789, 987654321, 5)
    got = ctx.poly_eval(coeffs, 1, n, 2, _mont3(p).reshape(1, 3))
    for c in range(2):
        assert np.array_equal(got[c, 0], orc.horner(coeffs[c], 1, _mont3(p)))


@pytest.mark.parametrize("fq", [3])
def test_deep_composition_matches_coefficient_form(ctx, orc, fq):
    """brainfuck-like shape: Fp base columns, Fq3 extension + composition columns, offsets {0, 1}."""
    torch = pytest.importorskip("torch")
    log_n, log_b = 9, 2
    n, N = 1 << log_n, 1 << (log_n + log_b)
    nbase, next_, ncomp = 3, 2, 2
    rng = random.Random(42)
    rq = lambda: tuple(rng.randrange(P) for _ in range(3))
    base_polys = orc.rand_matrix(nbase, n, 1, seed=1)
    ext_polys = orc.rand_matrix(next_, n, 3, seed=2)
    comp_polys = orc.rand_matrix(ncomp, n, 3, seed=3)
    trace_arguments = [(0, 0), (0, 1), (1, 0), (2, 1), (3, 0), (3, 1), (4, 0)]
    z = rq()
    z_points, z_m = deep.ood_points(z, log_n, [o for _, o in trace_arguments], ncomp)
    # ---- OOD evaluations on the GPU (get_ood_evals)
    pts = np.stack([_mont3(z_points[0]), _mont3(z_points[1])])
    base_ood = ctx.poly_eval(base_polys, 1, n, nbase, pts)
    ext_ood = ctx.poly_eval(ext_polys, 3, n, next_, pts)
    comp_ood = ctx.poly_eval(comp_polys, 3, n, ncomp, _mont3(z_m).reshape(1, 3))
    canon = lambda w: tuple(ms.from_mont(int(x)) for x in w)
    trace_oods = []
    for col, off in trace_arguments:
        w = base_ood[col, off] if col < nbase else ext_ood[col - nbase, off]
        want = orc.horner(base_polys[col] if col < nbase else ext_polys[col - nbase], 1 if col < nbase else 3,
                          _mont3(z_points[off]))
        assert np.array_equal(w, want)
        trace_oods.append(canon(w))
    comp_oods = [canon(comp_ood[j, 0]) for j in range(ncomp)]
    trace_alphas = [rq() for _ in trace_arguments]
    comp_alphas = [rq() for _ in range(ncomp)]
    d_alpha, d_beta = rq(), rq()

    # ---- reference path on the oracle: coefficient form, then LDE (bit-reversed)
    def lift_fp(col):
        out = np.zeros(3 * n, dtype=np.uint64)
        out[0::3] = col
        return out
    quotients = []
    for j in range(ncomp):
        quotients.append(orc.divide_out_points(comp_polys[j], _mont3(z_m), _mont3(comp_alphas[j])))
    for col in range(nbase + next_):
        zs = [z_points[o] for (c, o) in trace_arguments if c == col]
        cs = [a for (c, o), a in zip(trace_arguments, trace_alphas) if c == col]
        coeffs = lift_fp(base_polys[col]) if col < nbase else ext_polys[col - nbase]
        quotients.append(orc.divide_out_points(coeffs, np.concatenate([_mont3(p) for p in zs]),
                                               np.concatenate([_mont3(a) for a in cs])))
    combined = orc.sum_columns(np.stack(quotients), 3)
    combined = orc.degree_adjust(combined, _mont3(d_alpha), _mont3(d_beta))
    want = orc.lde(combined.reshape(1, -1), 3, log_n, log_b, orc.generator(), bitrev=True)[0]

    # ---- B200 path: pointwise over the resident bit-reversed LDEs
    gen = orc.generator()
    lde_of = lambda m, lanes: torch.from_numpy(orc.lde(m, lanes, log_n, log_b, gen, True).view(np.int64)).cuda()
    base_lde, ext_lde, comp_lde = lde_of(base_polys, 1), lde_of(ext_polys, 3), lde_of(comp_polys, 3)
    ex = deep.deep_expression(trace_arguments, nbase, next_, ncomp, z_points, z_m, trace_oods, comp_oods,
                              trace_alphas, comp_alphas, d_alpha, d_beta)
    prog = E.compile_program(ex, nbase, log_ce=log_n + log_b)
    out = torch.empty(3 * N, dtype=torch.int64, device="cuda")
    cols = [base_lde[c] for c in range(nbase)] + [ext_lde[c] for c in range(next_)] + [comp_lde[c] for c in range(ncomp)]
    ctx.eval_constraints_ptrs(prog, out, log_n + log_b, cols, [0] * nbase + [1] * (next_ + ncomp), fq_field=3,
                              offset=ms.GENERATOR, trace_bitrev=True, out_bitrev=True)
    ctx.sync()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), want)
