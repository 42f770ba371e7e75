"""helpers shared by tests/test_expr_compiler.py and tests/test_cpp_host.py: a big-integer interpreter of the evaluator's
instruction set (csrc/eval.cu) and a direct recursive evaluation of an expression DAG"""
import sys

from ministark_b200 import expr as E

P = E.P
R = 2**64
RINV = pow(R, -1, P)


def run_program(prog, x, cols, col_is_q, row, m):
    """big-int interpreter of the 4-word instructions; returns the stored Fq element (3-tuple of canonical ints)"""
    regs = {}
    out = None
    for op_w, d, a, b in prog.code.tolist():
        op, qa, qb = op_w & 0xFF, (op_w >> 8) & 1, (op_w >> 9) & 1
        if op == E.OP_X:
            regs[d] = (x, 0, 0)
        elif op == E.OP_CONST:
            regs[d] = tuple(int(w) * RINV % P for w in prog.consts[a])
        elif op == E.OP_TRACE:
            v = cols[a][(row + b) % m]
            assert bool(qa) == bool(col_is_q[a])
            regs[d] = tuple(v) if qa else (v, 0, 0)
        elif op == E.OP_NEG:
            regs[d] = E.q_neg(regs[a])
        elif op == E.OP_ADD:
            regs[d] = E.q_add(regs[a], regs[b])
        elif op == E.OP_SUB:
            regs[d] = E.q_add(regs[a], E.q_neg(regs[b]))
        elif op == E.OP_PERIODIC:
            v = cols[a][row % (1 << b)]          # the table rides in the column list after the trace columns
            regs[d] = tuple(v) if qa else (v, 0, 0)
        elif op == E.OP_MUL:
            regs[d] = E.q_mul(regs[a], regs[b])
        elif op == E.OP_INV:
            regs[d] = E.q_inv(regs[a]) if any(regs[a]) else (0, 0, 0)
        elif op == E.OP_POW:
            regs[d] = E.q_pow(regs[a], b)
        elif op == E.OP_STORE:
            out = regs[a]
        else:
            raise AssertionError(op)
        assert d < E.MAX_REGS
    return out


def periodic_value(coeffs, interval, x, n):
    """P(x^(n / interval)) by Horner, the verifier's formula (src/verifier.rs:221-230)"""
    y = pow(x, n // interval, P)
    acc = (0, 0, 0)
    for c in reversed(coeffs):
        c = c if isinstance(c, tuple) else (c, 0, 0)
        acc = E.q_add(E.q_mul(acc, (y, 0, 0)), c)
    return acc


def direct(expr, x, cols, row, m, challenges=(), hints=(), ccoefs=(), lde_step=1):
    memo = {}

    def ev(e):
        if id(e) in memo:
            return memo[id(e)]
        k, a = e.kind, e.args
        if k == "x":
            v = (x, 0, 0)
        elif k == "const":
            v = tuple(a[0])
        elif k == "chal":
            v = E._q(challenges[a[0]])
        elif k == "hint":
            v = E._q(hints[a[0]])
        elif k == "ccoef":
            v = E._q(ccoefs[a[0]])
        elif k == "trace":
            t = cols[a[0]][(row + lde_step * a[1]) % m]
            v = tuple(t) if isinstance(t, tuple) else (t, 0, 0)
        elif k == "periodic":
            v = periodic_value(a[0], a[1], x, m // lde_step)
        elif k == "neg":
            v = E.q_neg(ev(a[0]))
        elif k == "add":
            v = E.q_add(ev(a[0]), ev(a[1]))
        elif k == "mul":
            v = E.q_mul(ev(a[0]), ev(a[1]))
        elif k == "div":
            den = ev(a[1])
            v = E.q_mul(ev(a[0]), E.q_inv(den) if any(den) else (0, 0, 0))
        elif k == "pow":
            v = E.q_pow(ev(a[0]), a[1])
        else:
            raise AssertionError(k)
        memo[id(e)] = v
        return v

    import sys
    sys.setrecursionlimit(20000)
    return ev(expr)


def random_columns(rng, nbase, next_, m):
    cols = [[rng.randrange(P) for _ in range(m)] for _ in range(nbase)]
    cols += [[tuple(rng.randrange(P) for _ in range(3)) for _ in range(m)] for _ in range(next_)]
    return cols, [False] * nbase + [True] * next_


