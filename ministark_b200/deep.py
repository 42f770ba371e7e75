"""deep.py — DEEP composition (DeepPolyComposer, src/composer.rs:43-188) on resident data.

Reference: `get_ood_evals` Horner-evaluates every polynomial at the out-of-domain points, then
`into_deep_poly` builds  sum_k alpha_k (P(X) - P(z_k)) / (X - z_k)  in COEFFICIENT form by synthetic
division (a sequential recurrence per column, src/utils.rs:154-175), sums the columns, applies the degree
adjustment (alpha + beta X), and the prover then LDE-s the result in bit-reversed order
(src/prover.rs:146-148).

Here:
  * OOD evaluations: `Context.poly_eval` (csrc/deep.cu), a parallel reduction per (column, point);
  * the DEEP polynomial is evaluated directly over the LDE domain, pointwise, from the LDE columns that are
    already resident (bit-reversed): for an LDE point x
        DEEP(x) = (d_alpha + d_beta x) * [ sum_j a'_j (H_j(x) - H_j(z^m)) / (x - z^m)
                                          + sum_(c,o) a_(c,o) (T_c(x) - T_c(z g^o)) / (x - z g^o) ]
    P(X) - P(z) is divisible by X - z, so this is the value of the reference's coefficient-form polynomial
    at x — bit for bit — without the synthetic division, the column sum and the final LDE.  The expression
    is built with ministark_b200.expr and run by the fused evaluator (csrc/eval.cu) with
    trace_bitrev + out_bitrev, so the result comes out in the bit-reversed order FRI consumes.
"""
from . import expr as E

P = E.P


def ood_points(z, log_n, offsets, num_composition_cols):
    """z * g^offset for every distinct trace offset (g = trace-domain generator, g^-1 for negative offsets,
    src/composer.rs:63-66), and z^m for the composition columns (src/composer.rs:78)."""
    g = pow(pow(7, (P - 1) >> 32, P), 1 << (32 - log_n), P)
    g_inv = pow(g, -1, P)
    pts = {}
    for o in sorted(set(offsets)):
        s = pow(g if o >= 0 else g_inv, abs(o), P)
        pts[o] = tuple(c * s % P for c in z)
    z_m = E.q_pow(tuple(z), num_composition_cols)
    return pts, z_m


def deep_expression(trace_arguments, num_base_cols, num_ext_cols, num_composition_cols, z_points, z_m,
                    trace_oods, composition_oods, trace_alphas, composition_alphas, degree_alpha, degree_beta):
    """Expr over LDE columns ordered [base..., ext..., composition...] (all offsets 0: evaluation form).

    trace_arguments: list of (column, row offset) as Air::trace_arguments();  trace_oods / trace_alphas follow
    that order; z_points: {offset: z*g^offset}; composition_* follow the composition column order.
    All field values are canonical 3-tuples."""
    x = E.X()
    inv_cache = {}

    def inv_x_minus(pt):
        if pt not in inv_cache:
            inv_cache[pt] = E.Constant(1) / (x - E.Constant(pt))
        return inv_cache[pt]

    total = None
    ncols_trace = num_base_cols + num_ext_cols
    for j in range(num_composition_cols):
        term = (E.Trace(ncols_trace + j, 0) - E.Constant(composition_oods[j])) * inv_x_minus(tuple(z_m)) \
            * E.Constant(composition_alphas[j])
        total = term if total is None else total + term
    for (col, off), ood, alpha in zip(trace_arguments, trace_oods, trace_alphas):
        term = (E.Trace(col, 0) - E.Constant(ood)) * inv_x_minus(tuple(z_points[off])) * E.Constant(alpha)
        total = term if total is None else total + term
    return total * (E.Constant(degree_alpha) + x * E.Constant(degree_beta))


def deep_expression_symbolic(trace_arguments, num_base_cols, num_ext_cols, num_composition_cols):
    """The same expression with every per-proof value (out-of-domain points and values, alphas, degree coefficients)
    as a Hint placeholder, so that it is compiled ONCE per AIR (expr.compile_program(..., symbolic=True)) and bound per
    proof (Program.bind(hints=deep_hint_values(...))).  Returns (expr, keys); keys[i] names hint i."""
    keys, index = [], {}

    def H(key):
        if key not in index:
            index[key] = len(keys)
            keys.append(key)
        return E.Hint(index[key])

    x = E.X()
    inv_cache = {}

    def inv_x_minus(key):
        if key not in inv_cache:
            inv_cache[key] = E.Constant(1) / (x - H(key))
        return inv_cache[key]

    total = None
    ncols_trace = num_base_cols + num_ext_cols
    for j in range(num_composition_cols):
        term = (E.Trace(ncols_trace + j, 0) - H(("cood", j))) * inv_x_minus(("zm",)) * H(("calpha", j))
        total = term if total is None else total + term
    for i, (col, off) in enumerate(trace_arguments):
        term = (E.Trace(col, 0) - H(("tood", i))) * inv_x_minus(("zpt", off)) * H(("talpha", i))
        total = term if total is None else total + term
    return total * (H(("dalpha",)) + x * H(("dbeta",))), keys


def deep_hint_values(keys, z_points, z_m, trace_oods, composition_oods, trace_alphas, composition_alphas, degree_alpha, degree_beta):
    table = {"zm": lambda: z_m, "dalpha": lambda: degree_alpha, "dbeta": lambda: degree_beta,
             "zpt": lambda off: z_points[off], "tood": lambda i: trace_oods[i], "cood": lambda j: composition_oods[j],
             "talpha": lambda i: trace_alphas[i], "calpha": lambda j: composition_alphas[j]}
    return [tuple(table[k[0]](*k[1:])) for k in keys]
