"""eval_oracle.py — CPU oracle for constraint evaluation.  TEST INFRASTRUCTURE ONLY.

Restates AirConfig::eval_constraint -> eval_cpu::eval (src/air.rs:86-128, src/eval_cpu.rs:33-150,
258-493) as a memoised tree walk over whole columns with the C oracle's vector field ops
(the reference walks the same DAG per 512-element chunk; the field elements are identical).

Expression exchange format: nested tuples
    ('x',) | ('const', (c0,c1,c2), is_ext) | ('chal', i) | ('hint', i) | ('trace', col, row_offset)
    ('periodic', coeffs, interval_size)            src/constraints.rs:107-146, src/eval_cpu.rs:176-256
    ('neg', a) | ('add', a, b) | ('mul', a, b) | ('div', a, b) | ('pow', a, e)
Columns: natural-order evaluations over the ce coset (what bit_reverse_ce_trace hands to
eval_constraint, src/prover.rs:86-91).
"""
import numpy as np

from . import oracle as orc
from . import pyspec as S


def x_lde(log_m, offset_mont):
    """ce_lde_xs.elements(): offset * g^i, i natural (src/prover.rs:95)."""
    m = 1 << log_m
    x = np.empty(m, dtype=np.uint64)
    x[:] = offset_mont
    # distribute powers: x[i] *= g^i  via an NTT-free running product done in C (pointwise with shift)
    g = orc.root_of_unity(log_m)
    pw = np.empty(m, dtype=np.uint64)
    acc = orc.ONE
    # doubling construction: pw[0:k] known -> pw[k:2k] = pw[0:k] * g^k
    pw[0] = orc.ONE
    k = 1
    while k < m:
        gk = np.array([orc.fp_pow(g, k)], dtype=np.uint64)
        pw[k:2 * k] = orc.pointwise_const("mul", pw[0:k].copy(), 1, gk, 1)
        k *= 2
    return orc.pointwise("mul", x, 1, pw, 1)


def evaluate(expr, log_m, offset_mont, base_cols, ext_cols=None, fq_lanes=1, challenges=(), hints=(), lde_step=1):
    """returns the M x fq_lanes result column (numpy uint64, Montgomery words), natural order"""
    m = 1 << log_m
    nbase = 0 if base_cols is None else base_cols.shape[0]
    xs = x_lde(log_m, offset_mont)
    memo = {}

    def mont_const(v, lanes):
        return np.array([S.to_mont(int(c)) for c in v[:lanes]], dtype=np.uint64)

    def lanes_of(val):
        return val[1]

    def as_array(val):
        """(array-or-const, lanes, is_const) -> full array"""
        arr, lanes, is_const = val
        if not is_const:
            return arr
        return orc.pointwise_const("fill", None, 1, arr, lanes, n=m, dfield=lanes)

    def walk(e):
        if id(e) in memo:
            return memo[id(e)]
        k = e[0]
        if k == 'x':
            r = (xs, 1, False)
        elif k == 'const':
            lanes = fq_lanes if e[2] else 1
            r = (mont_const(e[1], lanes), lanes, True)
        elif k == 'chal':
            v = challenges[e[1]]
            v = (v, 0, 0) if isinstance(v, int) else v
            r = (mont_const(v, fq_lanes), fq_lanes, True)
        elif k == 'hint':
            v = hints[e[1]]
            v = (v, 0, 0) if isinstance(v, int) else v
            r = (mont_const(v, fq_lanes), fq_lanes, True)
        elif k == 'trace':
            col, off = e[1], e[2]
            shift = (lde_step * off) % m
            if col < nbase:
                src, lanes = base_cols[col], 1
            else:
                src, lanes = ext_cols[col - nbase], fq_lanes
            r = (np.roll(src.reshape(m, lanes), -shift, axis=0).reshape(-1).copy(), lanes, False)
        elif k == 'periodic':
            # eval_periodic_column (src/eval_cpu.rs:234-256) by the definition: P(y) at y = (offset * g_m^i)^(n / interval),
            # one evaluation per point of the domain of size interval * lde_step, repeated along the ce domain
            coeffs, interval = e[1], e[2]
            n = m // lde_step
            lanes = fq_lanes if any(isinstance(c, tuple) for c in coeffs) else 1
            period = interval * lde_step
            off = S.from_mont(int(offset_mont))
            g = S.root_of_unity(log_m)
            table = []
            for i in range(period):
                y = pow(off * pow(g, i, S.P) % S.P, n // interval, S.P)
                cs = [c if isinstance(c, tuple) else (c, 0, 0) for c in coeffs]
                acc = [0, 0, 0]
                for c in reversed(cs):
                    acc = [(acc[w] * y + c[w]) % S.P for w in range(3)]
                table.append([S.to_mont(acc[w]) for w in range(lanes)])
            tab = np.array(table, dtype=np.uint64).reshape(period, lanes)
            r = (np.tile(tab, (m // period, 1)).reshape(-1).copy(), lanes, False)
        elif k == 'neg':
            a = walk(e[1])
            r = (orc.pointwise("neg", as_array(a), a[1]), a[1], False)
        elif k in ('add', 'mul'):
            a, b = walk(e[1]), walk(e[2])
            lanes = max(a[1], b[1])
            r = (orc.pointwise(k, as_array(a), a[1], as_array(b), b[1], dfield=lanes), lanes, False)
        elif k == 'div':
            a, b = walk(e[1]), walk(e[2])
            lanes = max(a[1], b[1])
            binv = orc.pointwise("inv", as_array(b), b[1])
            r = (orc.pointwise("mul", as_array(a), a[1], binv, b[1], dfield=lanes), lanes, False)
        elif k == 'pow':
            a = walk(e[1])
            r = (orc.pointwise("exp", as_array(a), a[1], exponent=e[2]), a[1], False)
        else:
            raise ValueError(k)
        memo[id(e)] = r
        return r

    res = walk(expr)
    arr = as_array(res)
    if res[1] < fq_lanes:
        arr = orc.pointwise("convert", arr, 1, dfield=fq_lanes)
    return arr
