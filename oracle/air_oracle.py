"""air_oracle.py — the AIR bookkeeping of the reference restated INDEPENDENTLY of ministark_b200/air.py.  TEST INFRASTRUCTURE.

    degree rules of a constraint expression      src/constraints.rs:404-455   (numerator, denominator) degrees
    Constraint::blowup_factor                     src/constraints.rs:142-146,340-347
    AirConfig::composition_constraint            src/air.rs:50-82             sum_i c_i * (x^adj_i * alpha_i + beta_i)
    Air::ce_blowup_factor, trace_arguments, coefficient / challenge counts    src/air.rs:142-160,210-246

Works on the nested-tuple exchange format of oracle/eval_oracle.py (plus ('ccoef', i) for composition coefficients):
the prover under test and this restatement share only the LIST OF CONSTRAINTS, handed over as tuples; for examples/fib the
constraints themselves are restated below from examples/fib/main.rs:74-143, so that AIR has no shared statement at all.
`OracleAir` offers the attributes oracle/stark_oracle.py's prover and verifier read from an AIR description.
"""
from . import pyspec as S

P = S.P


def _walk(expr):
    """post-order list of the distinct nodes of a tuple DAG"""
    order, seen, stack = [], set(), [(expr, False)]
    while stack:
        node, done = stack.pop()
        if done:
            order.append(node)
            continue
        if id(node) in seen:
            continue
        seen.add(id(node))
        stack.append((node, True))
        for a in node[1:]:
            if isinstance(a, tuple) and a and isinstance(a[0], str) and id(a) not in seen:
                stack.append((a, False))
    return order


def degree(expr, trace_degree):
    """src/constraints.rs:404-455: X has degree 1, a trace cell trace_degree, constants / challenges / hints 0; Add cross-multiplies
    (max(an + bd, bn + ad), ad + bd); Mul adds, Div swaps, Pow scales, Neg keeps; a periodic column (len - 1) * n / interval"""
    d = {}
    for nd in _walk(expr):
        k = nd[0]
        if k in ("const", "chal", "hint", "ccoef"):
            v = (0, 0)
        elif k == "x":
            v = (1, 0)
        elif k == "trace":
            v = (trace_degree, 0)
        elif k == "periodic":
            v = ((len(nd[1]) - 1) * ((trace_degree + 1) // nd[2]), 0)
        elif k == "neg":
            v = d[id(nd[1])]
        elif k == "add":
            (an, ad), (bn, bd) = d[id(nd[1])], d[id(nd[2])]
            v = (max(an + bd, bn + ad), ad + bd)
        elif k == "mul":
            (an, ad), (bn, bd) = d[id(nd[1])], d[id(nd[2])]
            v = (an + bn, ad + bd)
        elif k == "div":
            (an, ad), (bn, bd) = d[id(nd[1])], d[id(nd[2])]
            v = (an + bd, ad + bn)
        elif k == "pow":
            n_, d_ = d[id(nd[1])]
            v = (n_ * nd[2], d_ * nd[2])
        else:
            raise ValueError(k)
        d[id(nd)] = v
    return d[id(expr)]


def _ceil_pow2(v):
    return 1 if v == 0 else (v if v & (v - 1) == 0 else 1 << v.bit_length())


def blowup_factor(expr, trace_len):
    num, den = degree(expr, trace_len - 1)
    return _ceil_pow2(max(num - den, 0)) // (trace_len - 1)


def _leaves(expr, kind):
    return {nd[1:] for nd in _walk(expr) if nd[0] == kind}


def composition_constraint(constraints, trace_len):
    """src/air.rs:50-82 — returns (expression tuple, ce_blowup_factor)"""
    ce = max(blowup_factor(c, trace_len) for c in constraints)
    composition_degree = trace_len * ce - 1
    x = ("x",)
    total = None
    for i, c in enumerate(constraints):
        num, den = degree(c, trace_len - 1)
        adj = composition_degree - (num - den)
        assert adj >= 0
        term = ("mul", c, ("add", ("mul", ("pow", x, adj), ("ccoef", 2 * i)), ("ccoef", 2 * i + 1)))
        total = term if total is None else ("add", total, term)
    return total, blowup_factor(total, trace_len)


def structurally_equal(a, b):
    """tuple DAGs equal as TREES (sharing is irrelevant), iteratively"""
    stack, seen = [(a, b)], set()
    while stack:
        x, y = stack.pop()
        if (id(x), id(y)) in seen:
            continue
        seen.add((id(x), id(y)))
        if isinstance(x, tuple) and x and isinstance(x[0], str):
            if not (isinstance(y, tuple) and len(x) == len(y) and x[0] == y[0]):
                return False
            stack.extend(zip(x[1:], y[1:]))
        elif x != y:
            return False
    return True


class OracleAir:
    def __init__(self, constraints, trace_len, public_inputs, options5, gen_hints, num_base_columns):
        self.constraints, self.trace_len, self.public_inputs, self.options5 = list(constraints), trace_len, public_inputs, options5
        self._gen_hints, self.num_base_columns = gen_hints, num_base_columns
        self.composition_tuple, self.ce_blowup_factor = composition_constraint(self.constraints, trace_len)
        assert self.ce_blowup_factor <= options5[1]

    def num_challenges(self):
        idx = [a[0] for c in self.constraints for a in _leaves(c, "chal")]
        return max(idx) + 1 if idx else 0

    def num_composition_constraint_coeffs(self):
        return 2 * len(self.constraints)

    def trace_arguments(self):
        args = set()
        for c in self.constraints:
            args |= _leaves(c, "trace")
        return sorted(args)

    def gen_hints(self, challenges):
        return self._gen_hints(self.trace_len, self.public_inputs, challenges)

    # the two forms oracle/stark_oracle.py asks for (objects with .to_tuple())
    class _T:
        def __init__(self, t):
            self.t = t

        def to_tuple(self):
            return self.t

    @property
    def composition_constraint(self):
        return OracleAir._T(self.composition_tuple)

    def substitute_composition_coeffs(self, coeffs):
        memo = {}
        for nd in _walk(self.composition_tuple):
            if nd[0] == "ccoef":
                v = coeffs[nd[1]]
                memo[id(nd)] = ("const", tuple(v), True) if isinstance(v, (tuple, list)) else ("const", (int(v) % P, 0, 0), True)
            else:
                memo[id(nd)] = (nd[0],) + tuple(memo[id(a)] if isinstance(a, tuple) and a and isinstance(a[0], str) else a for a in nd[1:])
        return OracleAir._T(memo[id(self.composition_tuple)])


# ---- examples/fib stated from examples/fib/main.rs:74-143 (8 columns of consecutive terms v_k = v_(k-2) * v_(k-1)) ---------
def fib_constraints(trace_len):
    g = S.root_of_unity(trace_len.bit_length() - 1)
    const = lambda v: ("const", (v % P, 0, 0), False)
    sub = lambda a, b: ("add", a, ("neg", b))                 # the reference's Sub is add(neg)
    x = ("x",)
    first, last, one = const(1), const(pow(g, trace_len - 1, P)), const(1)
    # main.rs:88-97 keeps the boundary values symbolic: v1 = v0 + v0, v2 = v1 * v0, v_k = v_(k-2) * v_(k-1)
    v = [one, ("add", one, one)]
    v.append(("mul", v[1], v[0]))
    for i in range(3, 8):
        v.append(("mul", v[i - 2], v[i - 1]))
    out = [("div", sub(("trace", i, 0), v[i]), sub(x, first)) for i in range(8)]
    out.append(("div", sub(("trace", 7, 0), ("hint", 0)), sub(x, last)))
    T = lambda c, o: ("trace", c, o)
    steps = [(0, (6, 0), (7, 0)), (1, (7, 0), (0, 1)), (2, (0, 1), (1, 1)), (3, (1, 1), (2, 1)), (4, (2, 1), (3, 1)),
             (5, (3, 1), (4, 1)), (6, (4, 1), (5, 1)), (7, (5, 1), (6, 1))]
    zerofier = ("div", sub(x, last), sub(("pow", x, trace_len), one))
    for col, a, b in steps:
        out.append(("mul", sub(T(col, 1), ("mul", T(*a), T(*b))), zerofier))
    return out


def fib_hints(trace_len, claimed, challenges):
    return [claimed]
