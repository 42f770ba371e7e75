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
    """The same polynomial with every per-proof value as a Hint placeholder, so that it is compiled ONCE per AIR
    (expr.compile_program(..., symbolic=True)) and bound per proof (Program.bind(hints=deep_hint_values(...))), and with
    the terms GROUPED by their out-of-domain point:

        sum_j a_j (P_j(x) - P_j(z_k)) / (x - z_k)  =  ( sum_j a_j P_j(x)  -  K_k ) / (x - z_k),   K_k = sum_j a_j P_j(z_k)

    K_k is one constant per distinct point (z^m, and z g^o per trace offset o), computed on the host from the OOD values
    the channel already holds.  Per LDE point this is one multiplication by a_j per column (Fq x Fp for a base-field
    column: 3 base-field products) and ONE Fq x Fq product per distinct point, instead of a subtraction and two Fq x Fq
    products per column — the same field element, hence the same words.  Returns (expr, keys); keys[i] names hint i."""
    keys, index = [], {}

    def H(key):
        if key not in index:
            index[key] = len(keys)
            keys.append(key)
        return E.Hint(index[key])

    x = E.X()
    ncols_trace = num_base_cols + num_ext_cols
    groups = [(("zm",), ("kzm",), [(ncols_trace + j, ("calpha", j)) for j in range(num_composition_cols)])]
    for off in sorted(set(o for _, o in trace_arguments)):
        groups.append((("zpt", off), ("kz", off), [(col, ("talpha", i)) for i, (col, o) in enumerate(trace_arguments) if o == off]))
    total = None
    for point, konst, members in groups:
        if not members:
            continue
        acc = None
        for col, alpha in members:
            term = E.Trace(col, 0) * H(alpha)
            acc = term if acc is None else acc + term
        term = (acc - H(konst)) * (E.Constant(1) / (x - H(point)))
        total = term if total is None else total + term
    return total * (H(("dalpha",)) + x * H(("dbeta",))), keys


def deep_hint_values(keys, z_points, z_m, trace_oods, composition_oods, trace_alphas, composition_alphas, degree_alpha, degree_beta,
                     trace_arguments=None):
    """values of the hints `keys` names.  trace_arguments (the list the expression was built from) is needed for the
    per-offset constants K_o = sum over the arguments with offset o of alpha * ood."""
    def dot(pairs):
        acc = (0, 0, 0)
        for a, v in pairs:
            acc = E.q_add(acc, E.q_mul(tuple(a), tuple(v)))
        return acc

    def kz(off):
        if trace_arguments is None:
            raise ValueError("deep_hint_values: trace_arguments is required for the grouped DEEP expression")
        return dot((trace_alphas[i], trace_oods[i]) for i, (_, o) in enumerate(trace_arguments) if o == off)

    table = {"zm": lambda: z_m, "dalpha": lambda: degree_alpha, "dbeta": lambda: degree_beta,
             "zpt": lambda off: z_points[off], "tood": lambda i: trace_oods[i], "cood": lambda j: composition_oods[j],
             "talpha": lambda i: trace_alphas[i], "calpha": lambda j: composition_alphas[j],
             "kzm": lambda: dot(zip(composition_alphas, composition_oods)), "kz": kz}
    return [tuple(table[k[0]](*k[1:])) for k in keys]
