"""GPU integration: every hot-path step of `default_prove` (src/prover.rs:46-155) chained on resident
data — base-trace commit, constraint evaluation, composition split + commit, OOD evaluations, DEEP
composition, FRI layers — against the same chain on the CPU oracle (which follows the reference's own
formulation: coefficient-form DEEP, apply_drp, Matrix::from_arrays + from_matrix).

The Fiat–Shamir channel is out of scope: both sides draw their "verifier randomness" from the
same deterministic stand-in coin (SHA-256 of the previous commitment), so the data flow and every
commitment are compared bit for bit while the protocol layer stays the reference's.
AIR: the wider variant of SURVEY.md §8d config 3 — T_k.next - prod_{j=1..5} T_{k-j} (degree 5 => ce_blowup 4),
Fq = Fp as in examples/fib."""
import hashlib

import numpy as np
import pytest

import ministark_b200 as ms
from ministark_b200 import deep
from ministark_b200 import expr as E

pytestmark = pytest.mark.gpu
P = ms.P


def coin(seed, k):
    """deterministic stand-in for PublicCoin::draw: k canonical field elements from SHA-256(seed || i)"""
    out = []
    i = 0
    while len(out) < k:
        v = int.from_bytes(hashlib.sha256(seed + i.to_bytes(4, "big")).digest()[:8], "big")
        i += 1
        if v < P:
            out.append(v)
    return out


def composition_expr(ncols, log_n, ce_blowup, coeffs):
    n = 1 << log_n
    g_inv = pow(pow(pow(7, (P - 1) >> 32, P), 1 << (32 - log_n), P), -1, P)
    x = E.X()
    num, den = x - E.Constant(g_inv), x ** n - E.Constant(1)
    comp_degree = n * ce_blowup - 1
    total = None
    for k in range(ncols):
        prod = E.Trace((k - 1) % ncols, 0)
        for j in range(2, 6):
            prod = prod * E.Trace((k - j) % ncols, 0)
        c = (E.Trace(k, 1) - prod) * num / den
        adj = comp_degree - ((5 * (n - 1) + 1) - n)
        term = c * (x ** adj * E.Constant(coeffs[2 * k]) + E.Constant(coeffs[2 * k + 1]))
        total = term if total is None else total + term
    return total


@pytest.mark.parametrize("log_n", [10])
def test_prover_hot_path_chain(orc, log_n):
    torch = pytest.importorskip("torch")
    from oracle import eval_oracle
    ctx = ms.Context(0)
    ncols, log_b, ce_blowup, log_ff = 8, 3, 4, 3
    log_ce, log_N = log_n + 2, log_n + log_b
    n, N, M = 1 << log_n, 1 << log_N, 1 << (log_n + 2)
    gen = orc.generator()
    dev = "cuda"
    t64 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(dev)
    u64 = lambda t: t.cpu().numpy().view(np.uint64)
    mont = lambda v: np.array([ms.to_mont(x) for x in v], dtype=np.uint64)

    # ---------------- 1. base trace commitment (prover.rs:46-55)
    trace = orc.rand_matrix(ncols, n, 1, seed=2025)
    d_trace = t64(trace)
    d_polys = torch.empty_like(d_trace)
    d_lde = torch.empty((ncols, N), dtype=torch.int64, device=dev)
    ctx.ntt_batch_to(d_trace, d_polys, ms.FP, log_n, ncols, inverse=True)
    ctx.lde_batch(d_polys, d_lde, ms.FP, log_n, log_b, ncols, offset=ms.GENERATOR, bitrev=True)
    root_base = ctx.merkle_commit(d_lde, ms.FP, N, ncols)
    o_polys = orc.ntt(trace, 1, log_n, inverse=True)
    o_lde = orc.lde(o_polys, 1, log_n, log_b, gen, True)
    assert root_base == orc.merkle_nodes(orc.hash_rows(o_lde, 1))[1].tobytes()

    # ---------------- 2. constraint evaluation over the ce domain (prover.rs:86-108)
    coeffs = coin(root_base, 2 * ncols)
    ex = composition_expr(ncols, log_n, ce_blowup, coeffs)
    prog = E.compile_program(ex, ncols, lde_step=ce_blowup, log_ce=log_ce)
    d_comp_evals = torch.empty(M, dtype=torch.int64, device=dev)
    ctx.eval_constraints(prog, d_comp_evals, log_ce, base_cols=d_lde, nbase=ncols, base_stride=N, fq_field=ms.FP,
                         offset=ms.GENERATOR, trace_bitrev=True)       # reads the bit-reversed LDE prefix in place
    ce_cols = np.stack([orc.bit_reverse(np.ascontiguousarray(o_lde[c][:M]), 1, log_ce) for c in range(ncols)])
    o_comp_evals = eval_oracle.evaluate(ex.to_tuple(), log_ce, gen, ce_cols, lde_step=ce_blowup)
    ctx.sync()
    assert np.array_equal(u64(d_comp_evals), o_comp_evals)

    # ---------------- 3. composition trace: iNTT over the ce coset, split, LDE, commit (prover.rs:110-125)
    ctx.ntt_batch(d_comp_evals, ms.FP, log_ce, 1, inverse=True, offset=ms.GENERATOR)
    d_comp_polys = torch.empty((ce_blowup, n), dtype=torch.int64, device=dev)
    ctx.matrix_from_rows(d_comp_evals, d_comp_polys, ms.FP, n, ce_blowup)       # column i = coefficients = i mod ce_blowup
    d_comp_lde = torch.empty((ce_blowup, N), dtype=torch.int64, device=dev)
    ctx.lde_batch(d_comp_polys, d_comp_lde, ms.FP, log_n, log_b, ce_blowup, offset=ms.GENERATOR, bitrev=True)
    root_comp = ctx.merkle_commit(d_comp_lde, ms.FP, N, ce_blowup)
    o_comp_poly = orc.ntt(o_comp_evals.reshape(1, -1), 1, log_ce, gen, inverse=True)[0]
    o_comp_polys = np.ascontiguousarray(o_comp_poly.reshape(n, ce_blowup).T)
    o_comp_lde = orc.lde(o_comp_polys, 1, log_n, log_b, gen, True)
    assert root_comp == orc.merkle_nodes(orc.hash_rows(o_comp_lde, 1))[1].tobytes()

    # ---------------- 4. OOD evaluations + DEEP composition (prover.rs:133-149, composer.rs)
    z = (coin(root_comp, 1)[0], 0, 0)
    trace_arguments = [(c, o) for c in range(ncols) for o in (0, 1)]
    z_points, z_m = deep.ood_points(z, log_n, [o for _, o in trace_arguments], ce_blowup)
    pts = np.stack([mont(z_points[0]), mont(z_points[1])])
    t_ood = ctx.poly_eval(d_polys, ms.FP, n, ncols, pts)
    c_ood = ctx.poly_eval(d_comp_polys, ms.FP, n, ce_blowup, mont(z_m).reshape(1, 3))
    for c in range(ncols):
        for o in (0, 1):
            assert np.array_equal(t_ood[c, o], orc.horner(o_polys[c], 1, mont(z_points[o])))
    canon = lambda w: tuple(ms.from_mont(int(x)) for x in w)
    trace_oods = [canon(t_ood[c, o]) for c, o in trace_arguments]
    comp_oods = [canon(c_ood[j, 0]) for j in range(ce_blowup)]
    rnd = coin(root_comp + b"deep", len(trace_arguments) + ce_blowup + 2)
    lift = lambda v: (v, 0, 0)
    trace_alphas = [lift(v) for v in rnd[:len(trace_arguments)]]
    comp_alphas = [lift(v) for v in rnd[len(trace_arguments):len(trace_arguments) + ce_blowup]]
    d_alpha, d_beta = lift(rnd[-2]), lift(rnd[-1])
    dex = deep.deep_expression(trace_arguments, ncols, 0, ce_blowup, z_points, z_m, trace_oods, comp_oods,
                               trace_alphas, comp_alphas, d_alpha, d_beta)
    dprog = E.compile_program(dex, ncols + ce_blowup, log_ce=log_N)      # Fq = Fp: every column is a base column
    d_deep = torch.empty(N, dtype=torch.int64, device=dev)
    cols = [d_lde[c] for c in range(ncols)] + [d_comp_lde[j] for j in range(ce_blowup)]
    ctx.eval_constraints_ptrs(dprog, d_deep, log_N, cols, [0] * len(cols), fq_field=ms.FP, offset=ms.GENERATOR,
                              trace_bitrev=True, out_bitrev=True)
    # oracle: the reference's coefficient form (src/composer.rs:89-188), then the LDE
    def lift_col(col):
        out = np.zeros(3 * n, dtype=np.uint64)
        out[0::3] = col
        return out
    quotients = [orc.divide_out_points(lift_col(o_comp_polys[j]), mont(z_m), mont(comp_alphas[j])) for j in range(ce_blowup)]
    for c in range(ncols):
        zs = np.concatenate([mont(z_points[o]) for (cc, o) in trace_arguments if cc == c])
        cs = np.concatenate([mont(a) for (cc, o), a in zip(trace_arguments, trace_alphas) if cc == c])
        quotients.append(orc.divide_out_points(lift_col(o_polys[c]), zs, cs))
    o_deep_poly = orc.degree_adjust(orc.sum_columns(np.stack(quotients), 3), mont(d_alpha), mont(d_beta))
    assert not o_deep_poly[1::3].any() and not o_deep_poly[2::3].any()      # Fq = Fp: stays in the base field
    o_deep = orc.lde(np.ascontiguousarray(o_deep_poly[0::3]).reshape(1, -1), 1, log_n, log_b, gen, True)[0]
    ctx.sync()
    assert np.array_equal(u64(d_deep), o_deep)

    # ---------------- 5. FRI layers (fri.rs:179-231): commit rows of ff evaluations, fold, repeat
    cur, want, ln = d_deep, o_deep, log_N
    seed = root_comp + b"fri"
    while ln - log_ff >= 5:
        nrows = 1 << (ln - log_ff)
        root_layer = ctx.merkle_commit_rows(cur, 1 << log_ff, nrows)
        layer_cols = np.ascontiguousarray(want.reshape(nrows, 1 << log_ff).T)          # Matrix::from_arrays
        assert root_layer == orc.merkle_nodes(orc.hash_rows(layer_cols, 1))[1].tobytes()
        alpha = mont(coin(seed + root_layer, 1))
        nxt = torch.empty(nrows, dtype=torch.int64, device=dev)
        ctx.fri_fold(cur, nxt, ms.FP, ln, log_ff, alpha)
        want = orc.fri_apply_drp(want, 1, ln, log_ff, alpha)
        ctx.sync()
        assert np.array_equal(u64(nxt), want)
        cur, ln, seed = nxt, ln - log_ff, root_layer
    # ---------------- 6. queries: rows of the resident LDE (trace.rs:115-157)
    positions = sorted(set(v % N for v in coin(seed + b"q", 12)))
    rows = ctx.gather_rows(d_lde, ms.FP, N, ncols, positions)
    assert np.array_equal(rows, o_lde[:, positions].T)
