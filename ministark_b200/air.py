"""air.py — host-side mirror of the reference's AIR description for the prover driver (prover.py).

    ProofOptions                      src/lib.rs:86-132
    AirConfig / Air                   src/air.rs:26-247
    constraint degree bookkeeping     src/constraints.rs:131-146,325-347,404-455
    composition constraint            src/air.rs:50-82

Constraints are `ministark_b200.expr.Expr` DAGs over the leaves X | Constant | Challenge(i) | Hint(i) |
Trace(column, offset) | Periodic(coeffs, interval_size) plus, inside the
composition constraint only, CompositionCoeff(i) — the verifier randomness that is substituted as
constants once the channel has produced it (src/air.rs:96-101).

This module is pure bookkeeping (a few hundred DAG nodes): no field data is touched here.
"""
from dataclasses import dataclass

from . import expr as E

P = E.P
TWO_ADIC_ROOT = pow(7, (P - 1) >> 32, P)        # arkworks' 2^32-th root of unity for Goldilocks
GENERATOR = 7                                    # Fp::GENERATOR, the LDE coset offset (src/air.rs:42-44)


def domain_generator(log_n):
    """Radix2EvaluationDomain::group_gen for size 2^log_n"""
    return pow(TWO_ADIC_ROOT, 1 << (32 - log_n), P)


@dataclass(frozen=True)
class ProofOptions:
    num_queries: int
    lde_blowup_factor: int
    grinding_factor: int
    fri_folding_factor: int
    fri_max_remainder_coeffs: int

    def __post_init__(self):
        # the reference's const asserts, src/lib.rs:109-114
        assert 1 <= self.num_queries <= 128
        b = self.lde_blowup_factor
        assert 1 <= b <= 128 and b & (b - 1) == 0
        assert self.grinding_factor <= 50
        assert self.fri_folding_factor in (2, 4, 8, 16)          # src/fri.rs:185-192

    def to_bytes(self):
        """derived CanonicalSerialize: the five u8 fields in declaration order"""
        return bytes([self.num_queries, self.lde_blowup_factor, self.grinding_factor, self.fri_folding_factor,
                      self.fri_max_remainder_coeffs])

    # FriOptions (src/fri.rs:30-69)
    def fri_num_layers(self, domain_size):
        k = 0
        while domain_size > self.fri_max_remainder_coeffs * self.lde_blowup_factor:
            domain_size //= self.fri_folding_factor
            k += 1
        return k

    def fri_remainder_size(self, domain_size):
        while domain_size > self.fri_max_remainder_coeffs * self.lde_blowup_factor:
            domain_size //= self.fri_folding_factor
        return domain_size


def CompositionCoeff(i):
    return E.Expr("ccoef", int(i))


def _ceil_power_of_two(v):
    """src/utils.rs:76-82 (0 is not a power of two and 0usize.next_power_of_two() == 1)"""
    if v == 0:
        return 1
    return v if v & (v - 1) == 0 else 1 << v.bit_length()


def degree(expr, trace_degree):
    """(numerator degree, denominator degree) by the reference's rules (src/constraints.rs:404-455): Add takes the
    cross-multiplied maximum and ADDS the denominators, Neg is the identity, Mul/Div/Pow as for rational functions."""
    memo = {}
    order, seen = [], set()
    stack = [(expr, False)]
    while stack:
        node, done = stack.pop()
        if done:
            order.append(node)
            continue
        if id(node) in seen:
            continue
        seen.add(id(node))
        stack.append((node, True))
        for a in node.args:
            if isinstance(a, E.Expr) and id(a) not in seen:
                stack.append((a, False))
    for nd in order:
        k, a = nd.kind, nd.args
        if k in ("const", "chal", "hint", "ccoef"):
            d = (0, 0)
        elif k == "trace":
            d = (trace_degree, 0)
        elif k == "x":
            d = (1, 0)
        elif k == "periodic":
            # PeriodicColumn::degree (src/constraints.rs:131-138): (len(coeffs) - 1) * (trace_len / interval_size)
            d = ((len(a[0]) - 1) * ((trace_degree + 1) // a[1]), 0)
        elif k == "neg":
            d = memo[id(a[0])]
        elif k == "add":
            (an, ad), (bn, bd) = memo[id(a[0])], memo[id(a[1])]
            d = (max(an + bd, bn + ad), ad + bd)
        elif k == "mul":
            (an, ad), (bn, bd) = memo[id(a[0])], memo[id(a[1])]
            d = (an + bn, ad + bd)
        elif k == "div":
            (an, ad), (bn, bd) = memo[id(a[0])], memo[id(a[1])]
            d = (an + bd, ad + bn)
        elif k == "pow":
            n, dd = memo[id(a[0])]
            d = (n * a[1], dd * a[1])
        else:
            raise ValueError(f"unsupported node {k}")
        memo[id(nd)] = d
    return memo[id(expr)]


def blowup_factor(expr, trace_len):
    """Constraint::blowup_factor (src/constraints.rs:142-146,340-347)"""
    trace_degree = trace_len - 1
    num, den = degree(expr, trace_degree)
    deg = max(num - den, 0)                                    # saturating_sub
    return _ceil_power_of_two(deg) // trace_degree


def _leaves(expr, kind):
    out, seen, stack = set(), set(), [expr]
    while stack:
        node = stack.pop()
        if id(node) in seen:
            continue
        seen.add(id(node))
        if node.kind == kind:
            out.add(node.args)
        stack.extend(a for a in node.args if isinstance(a, E.Expr))
    return out


class AirConfig:
    """Subclass and override, as with the reference's trait (src/air.rs:26-48).  Field values handed to and
    returned by the hooks are canonical integers (Fp) or 3-tuples of canonical integers (Fq3)."""
    NUM_BASE_COLUMNS = 0
    NUM_EXTENSION_COLUMNS = 0
    FQ_IS_FP = True              # type Fq = Fp (examples/fib) vs. Fq = Fq3 (examples/brainfuck)

    @staticmethod
    def constraints(trace_len):
        raise NotImplementedError

    @staticmethod
    def gen_hints(trace_len, public_inputs, challenges):
        return []

    @staticmethod
    def domain_offset():
        return GENERATOR


class Air:
    """Air::new (src/air.rs:142-160): constraints, the composition constraint and the ce blow-up factor."""

    def __init__(self, config, trace_len, public_inputs, options):
        assert trace_len & (trace_len - 1) == 0
        self.config, self.trace_len, self.public_inputs, self.options = config, trace_len, public_inputs, options
        self.log_n = trace_len.bit_length() - 1
        self.constraints = list(config.constraints(trace_len))
        # AirConfig::composition_constraint (src/air.rs:50-82)
        ce_blowup = max(blowup_factor(c, trace_len) for c in self.constraints)
        composition_degree = trace_len * ce_blowup - 1
        trace_degree = trace_len - 1
        x = E.X()
        total = None
        for i, c in enumerate(self.constraints):
            num, den = degree(c, trace_degree)
            evaluation_degree = num - den
            assert evaluation_degree <= composition_degree
            adj = composition_degree - evaluation_degree
            term = c * (x ** adj * CompositionCoeff(2 * i) + CompositionCoeff(2 * i + 1))
            total = term if total is None else total + term
        self.composition_constraint = total
        self.ce_blowup_factor = blowup_factor(total, trace_len)
        assert self.ce_blowup_factor <= options.lde_blowup_factor

    def lde_blowup_factor(self):
        return self.options.lde_blowup_factor

    # ---- compiled evaluator programs, shared by every proof of this AIR and trace length (the verifier randomness
    # enters through Program.bind, not through the instruction stream)
    def composition_program(self):
        if getattr(self, "_composition_program", None) is None:
            cfg = self.config
            log_ce = self.log_n + self.ce_blowup_factor.bit_length() - 1
            self._composition_program = E.compile_program(self.composition_constraint, cfg.NUM_BASE_COLUMNS,
                                                          lde_step=self.ce_blowup_factor, log_ce=log_ce, symbolic=True,
                                                          batch_inverses=True)   # zerofier denominators: never 0 on the LDE coset
        return self._composition_program

    def deep_program(self):
        if getattr(self, "_deep_program", None) is None:
            from . import deep
            cfg = self.config
            log_N = self.log_n + self.options.lde_blowup_factor.bit_length() - 1
            expr, keys = deep.deep_expression_symbolic(self.trace_arguments(), cfg.NUM_BASE_COLUMNS, cfg.NUM_EXTENSION_COLUMNS,
                                                       self.ce_blowup_factor)
            self._deep_program = (E.compile_program(expr, cfg.NUM_BASE_COLUMNS, log_ce=log_N, symbolic=True, max_live_leaves=8,
                                                    batch_inverses=True), keys)
        return self._deep_program

    # the three walks below depend on the constraints only: done once per Air (the provers copy a cached Air per proof,
    # and each walk of the brainfuck AIR costs ~0.5 ms of a 10 ms proof)
    def num_challenges(self):
        if getattr(self, "_num_challenges", None) is None:
            idx = [a[0] for c in self.constraints for a in _leaves(c, "chal")]
            self._num_challenges = max(idx) + 1 if idx else 0
        return self._num_challenges

    def num_composition_constraint_coeffs(self):
        if getattr(self, "_num_ccoefs", None) is None:
            idx = [a[0] for a in _leaves(self.composition_constraint, "ccoef")]
            self._num_ccoefs = max(idx) + 1 if idx else 0
        return self._num_ccoefs

    def gen_hints(self, challenges):
        return self.config.gen_hints(self.trace_len, self.public_inputs, challenges)

    def trace_arguments(self):
        """BTreeSet<(column, offset)> — sorted by column, then signed offset (src/air.rs:240-246)"""
        if getattr(self, "_trace_arguments", None) is None:
            args = set()
            for c in self.constraints:
                args |= _leaves(c, "trace")
            self._trace_arguments = sorted(args)
        return list(self._trace_arguments)

    def substitute_composition_coeffs(self, coeffs):
        """the map_leaves of AirConfig::eval_constraint (src/air.rs:96-101): CompositionCoeff(i) -> Constant"""
        memo = {}

        def sub(e):
            stack = [e]
            while stack:
                node = stack[-1]
                if id(node) in memo:
                    stack.pop()
                    continue
                todo = [a for a in node.args if isinstance(a, E.Expr) and id(a) not in memo]
                if todo:
                    stack.extend(todo)
                    continue
                if node.kind == "ccoef":
                    v = coeffs[node.args[0]]
                    memo[id(node)] = E.Constant(v) if isinstance(v, (tuple, list)) else E.Constant(v, ext=True)
                else:
                    memo[id(node)] = E.Expr(node.kind, *[memo[id(a)] if isinstance(a, E.Expr) else a for a in node.args])
                stack.pop()
            return memo[id(e)]

        return sub(self.composition_constraint)
