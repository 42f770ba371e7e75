"""CPU-only: the constraint-expression compiler (ministark_b200/expr.py) — operand ordering, register allocation with
leaf rematerialisation under the 48-register limit, constant folding, Div -> Mul/Inv rewriting, symbolic constants and
Program.bind — checked by EXECUTING the emitted programs with a big-integer interpreter of the evaluator's instruction
set (csrc/eval.cu) and comparing with a direct recursive evaluation of the expression DAG."""
import random

import numpy as np
import pytest

from ministark_b200 import expr as E
from ministark_b200.air import Air, CompositionCoeff, ProofOptions
from ministark_b200.examples import brainfuck as bf
from ministark_b200.examples import fib, perm

P = E.P
R = 2**64
RINV = pow(R, -1, P)


from tests_helpers_expr import direct, random_columns, run_program  # noqa: E402


@pytest.mark.parametrize("seed", range(6))
def test_random_expressions(seed):
    rng = random.Random(seed)
    nbase, next_, m = 4, 2, 8
    cols, is_q = random_columns(rng, nbase, next_, m)
    leaves = [E.X()] + [E.Trace(c, o) for c in range(nbase + next_) for o in (0, 1)] + [E.Constant(rng.randrange(P)) for _ in range(3)] \
        + [E.Constant(tuple(rng.randrange(P) for _ in range(3))), E.Challenge(0), E.Hint(0)]
    pool = list(leaves)
    for _ in range(60):
        a, b = rng.choice(pool), rng.choice(pool)
        pool.append(rng.choice([lambda: a + b, lambda: a - b, lambda: a * b, lambda: a / (b + E.Constant(1)), lambda: -a,
                                lambda: a ** rng.randrange(0, 5)])())
    expr = pool[-1] + pool[-2] * pool[-3] + pool[len(pool) // 2]
    ch, hi = [tuple(rng.randrange(P) for _ in range(3))], [tuple(rng.randrange(P) for _ in range(3))]
    plain = E.compile_program(expr, nbase, challenges=ch, hints=hi, lde_step=1, log_ce=3)
    sym = E.compile_program(expr, nbase, lde_step=1, log_ce=3, symbolic=True).bind(challenges=ch, hints=hi)
    for row in range(m):
        x = rng.randrange(1, P)
        want = direct(expr, x, cols, row, m, ch, hi)
        assert run_program(plain, x, cols, is_q, row, m) == want
        assert run_program(sym, x, cols, is_q, row, m) == want


@pytest.mark.parametrize("seed", range(6))
def test_batched_inverses_and_shared_powers(seed):
    """compile_program(batch_inverses=True): the inverses of one point through ONE inversion (Montgomery's trick), Fp and
    Fq denominators in separate batches, denominators that read the trace and an inverse under an inverse left alone; and
    x^(a n + b) rewritten to
    (x^n)^a * x^b.  Same values as the direct evaluation whenever no denominator vanishes (random operands)."""
    rng = random.Random(100 + seed)
    nbase, next_, m, lde_step = 4, 2, 32, 4
    n = m // lde_step
    cols, is_q = random_columns(rng, nbase, next_, m)
    x = E.X()
    T = E.Trace
    fp_dens = [x - E.Constant(rng.randrange(P)), x ** n - E.Constant(1), x * x + E.Constant(3), T(2, 0) * T(3, 1) + x]
    fq_dens = [x - E.Hint(0), x * E.Challenge(0) + E.Constant(9), T(5, 1) * x + T(4, 1)]
    expr = None
    for k, d in enumerate(fp_dens + fq_dens):
        term = (T(k % 4, 0) + E.Constant(k)) / d * (x ** ((3 + k) * n + k) * E.Challenge(0) + E.Hint(0))
        expr = term if expr is None else expr + term
    expr = expr + E.Constant(7) / (x + E.Constant(1) / (x - E.Constant(5)))          # an inverse under an inverse
    ch, hi = [tuple(rng.randrange(P) for _ in range(3))], [tuple(rng.randrange(P) for _ in range(3))]
    log_ce = m.bit_length() - 1
    plain = E.compile_program(expr, nbase, challenges=ch, hints=hi, lde_step=lde_step, log_ce=log_ce)
    batched = E.compile_program(expr, nbase, challenges=ch, hints=hi, lde_step=lde_step, log_ce=log_ce, batch_inverses=True)
    sym = E.compile_program(expr, nbase, lde_step=lde_step, log_ce=log_ce, symbolic=True, batch_inverses=True, max_live_leaves=4)
    sym = sym.bind(challenges=ch, hints=hi)
    count = lambda prog, op: int(sum(1 for ins in prog.code if ins[0] & 0xff == op))
    # 9 inversions; batched: one for the Fp denominators in x alone (with the inner 1/(x - 5)), one for the Fq ones, the
    # two denominators that read trace cells keep their own, and so does the outer inverse
    assert count(plain, E.OP_INV) == 9 and count(batched, E.OP_INV) == 5
    assert max(int(ins[3]) for ins in batched.code if ins[0] & 0xff == E.OP_POW) < 2 * n   # every x^(a n + b) was split: only x^n, a and b are left as exponents
    for row in range(0, m, 5):
        xv = rng.randrange(1, P)
        want = direct(expr, xv, cols, row, m, ch, hi, lde_step=lde_step)
        for prog in (plain, batched, sym):
            assert run_program(prog, xv, cols, is_q, row, m) == want


@pytest.mark.parametrize("which", ["fib", "perm", "brainfuck"])
def test_composition_programs_of_the_example_airs(which):
    """the real compositions: brainfuck's 48 constraints need leaf rematerialisation to fit 48 registers"""
    rng = random.Random(7)
    n = 16
    if which == "fib":
        cfg, pub, opts = fib.FibAirConfig, 5, fib.OPTIONS
    elif which == "perm":
        cfg, pub, opts = perm.PermAirConfig, [], ProofOptions(8, 8, 0, 4, 4)
    else:
        cfg, pub, opts, n = bf.BrainfuckAirConfig, bf.BrainfuckClaim("+.", b"", b"\x01"), bf.OPTIONS, 64
    air = Air(cfg, n, pub, opts)
    ce = air.ce_blowup_factor
    m = n * ce
    cols, is_q = random_columns(rng, cfg.NUM_BASE_COLUMNS, cfg.NUM_EXTENSION_COLUMNS, m)
    lift = (lambda v: v[0]) if cfg.FQ_IS_FP else (lambda v: v)
    ch = [lift(tuple(rng.randrange(P) for _ in range(3))) for _ in range(air.num_challenges())]
    hints = air.gen_hints(ch)
    cc = [lift(tuple(rng.randrange(P) for _ in range(3))) for _ in range(air.num_composition_constraint_coeffs())]
    prog = air.composition_program()
    assert prog.nregs <= E.MAX_REGS and len(prog.bindings) >= len(cc)
    bound = prog.bind(challenges=ch, hints=hints, ccoefs=cc)
    substituted = E.compile_program(air.substitute_composition_coeffs(cc), cfg.NUM_BASE_COLUMNS, challenges=ch, hints=hints,
                                    lde_step=ce, log_ce=m.bit_length() - 1)
    for row in (0, 1, m - 1, m // 2 + 3):
        x = rng.randrange(2, P)
        want = direct(air.composition_constraint, x, cols, row, m, ch, hints, cc, lde_step=ce)
        assert run_program(bound, x, cols, is_q, row, m) == want
        assert run_program(substituted, x, cols, is_q, row, m) == want


def test_register_pressure_is_reported():
    # a balanced tree of products of distinct trace cells cannot be evaluated in fewer registers than its depth allows:
    # interior temporaries are never evicted, so a wide enough expression must raise instead of miscompiling
    terms = [E.Trace(i % 7, 0) * E.Trace((i + 1) % 7, 1) + E.Constant(i + 1) for i in range(256)]
    while len(terms) > 1:
        terms = [terms[i] * terms[i + 1] for i in range(0, len(terms), 2)]
    prog = E.compile_program(terms[0], 7, log_ce=4)          # Sethi-Ullman order keeps this one small
    assert prog.nregs <= E.MAX_REGS
    rng = random.Random(1)
    cols, is_q = random_columns(rng, 7, 0, 16)
    assert run_program(prog, 3, cols, is_q, 5, 16) == direct(terms[0], 3, cols, 5, 16)


@pytest.mark.parametrize("which", ["fib", "perm"])
def test_deep_program_symbolic_equals_plain(which):
    from ministark_b200 import deep
    rng = random.Random(3)
    n = 16
    cfg, pub, opts = (fib.FibAirConfig, 5, fib.OPTIONS) if which == "fib" else (perm.PermAirConfig, [], ProofOptions(8, 8, 0, 4, 4))
    air = Air(cfg, n, pub, opts)
    nb, ne, nc = cfg.NUM_BASE_COLUMNS, cfg.NUM_EXTENSION_COLUMNS, air.ce_blowup_factor
    m = n * opts.lde_blowup_factor
    cols, is_q = random_columns(rng, nb, ne + nc, m)
    q3 = lambda: tuple(rng.randrange(P) for _ in range(3))
    args = air.trace_arguments()
    z = q3()
    z_points, z_m = deep.ood_points(z, air.log_n, sorted(set(o for _, o in args)), nc)
    toods, coods = [q3() for _ in args], [q3() for _ in range(nc)]
    talphas, calphas, da, db = [q3() for _ in args], [q3() for _ in range(nc)], q3(), q3()
    plain = E.compile_program(deep.deep_expression(args, nb, ne, nc, z_points, z_m, toods, coods, talphas, calphas, da, db), nb,
                              log_ce=m.bit_length() - 1)
    sym, keys = air.deep_program()
    bound = sym.bind(hints=deep.deep_hint_values(keys, z_points, z_m, toods, coods, talphas, calphas, da, db, trace_arguments=args))
    for row in (0, 7, m - 1):
        x = rng.randrange(2, P)
        assert run_program(bound, x, cols, is_q, row, m) == run_program(plain, x, cols, is_q, row, m)


def test_periodic_columns_compile_and_degree():
    """Periodic(coeffs, interval) (src/constraints.rs:107-146): the compiled program reads a table of interval * lde_step
    evaluations (eval_periodic_column, src/eval_cpu.rs:234-256) that repeats along the ce domain; the value must equal the
    verifier's P(x^(n / interval)) (src/verifier.rs:221-230) at every point x = offset * g^row"""
    from ministark_b200 import air as A
    from tests_helpers_expr import periodic_value
    rng = random.Random(3)
    n, lde_step = 8, 2
    m = n * lde_step
    log_m = m.bit_length() - 1
    g = A.domain_generator(log_m)
    pa = E.Periodic([3, 5], 4)                                                  # base-field coefficients
    pb = E.Periodic([(1, 2, 3), (4, 5, 6), (7, 8, 9), (10, 11, 12)], 8)         # extension coefficients
    cols, is_q = random_columns(rng, 2, 1, m)
    expr = (E.Trace(0, 1) - pa * E.Trace(1, 0)) * pb + E.Trace(2, 0) * pa * pa - E.X()
    prog = E.compile_program(expr, 2, lde_step=lde_step, log_ce=log_m, num_cols=3)
    assert [(slot, interval, q, ll) for slot, _, interval, q, ll in prog.periodic] == [(3, 4, False, 3), (4, 8, True, 4)] or \
        [(slot, interval, q, ll) for slot, _, interval, q, ll in prog.periodic] == [(3, 8, True, 4), (4, 4, False, 3)]
    xs = [A.GENERATOR * pow(g, i, P) % P for i in range(m)]
    tables, tq = [], []
    for _, coeffs, interval, q, log_len in prog.periodic:
        vals = [periodic_value(coeffs, interval, xs[i], n) for i in range(1 << log_len)]
        tables.append([v if q else v[0] for v in vals])
        tq.append(q)
        # the table really is periodic along the domain
        assert all(periodic_value(coeffs, interval, xs[i], n) == vals[i % (1 << log_len)] for i in range(m))
    for row in range(m):
        want = direct(expr, xs[row], cols, row, m, lde_step=lde_step)
        assert run_program(prog, xs[row], cols + tables, is_q + tq, row, m) == want
    # degree rule: (len(coeffs) - 1) * (trace_len / interval)
    assert A.degree(pa, n - 1) == (1 * (n // 4), 0)
    assert A.degree(pb * E.Trace(0, 0), n - 1) == (3 * (n // 8) + n - 1, 0)
    with pytest.raises(ValueError):
        E.Periodic([1, 2, 3], 4)
    with pytest.raises(ValueError):
        E.compile_program(pa, 2, lde_step=lde_step, log_ce=log_m)              # num_cols missing
