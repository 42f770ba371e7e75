"""Synthetic AIR of BASELINE config 3: workload definition used by bench.py and the tests.

32 (= ncols) transition constraints  T_k.next - T_{k-2}.curr * T_{k-1}.curr  (indices mod ncols; the
shape of examples/fib/main.rs:110-119), each multiplied by the transition zerofier
(X - g^-1) / (X^n - 1) and composed as  constraint * (X^adj * alpha_k + beta_k)  exactly like
AirConfig::composition_constraint (src/air.rs:50-82).  Degree rule (src/constraints.rs:340-347)
gives ce_blowup = 1 and adj = 0 for these degree-2 constraints, as for the real fib example.
"""
from . import FP, GENERATOR, P
from . import expr as E



def _coeff(k, which):
    """deterministic stand-ins for the verifier's composition coefficients (canonical ints)"""
    x = (0x9E3779B97F4A7C15 * (2 * k + which + 1)) % P
    return x


def composition(ncols, log_n, ce_blowup=1):
    n = 1 << log_n
    g = pow(pow(7, (P - 1) >> 32, P), 1 << (32 - log_n), P)
    g_inv = pow(g, -1, P)
    x = E.X()
    zerofier_num = x - E.Constant(g_inv)                 # all rows but the last one
    zerofier_den = x ** n - E.Constant(1)
    composition_degree = n * ce_blowup - 1
    total = None
    for k in range(ncols):
        c = E.Trace(k, 1) - E.Trace((k - 2) % ncols, 0) * E.Trace((k - 1) % ncols, 0)
        c = c * zerofier_num / zerofier_den
        evaluation_degree = (2 * (n - 1) + 1) - n        # numerator - denominator degree
        adj = composition_degree - evaluation_degree
        term = c * (x ** adj * E.Constant(_coeff(k, 0)) + E.Constant(_coeff(k, 1)))
        total = term if total is None else total + term
    return total


class GpuConstraintEval:
    """compiles the composition once; run() evaluates it straight off the bit-reversed LDE prefix."""

    def __init__(self, ctx, log_n, log_blowup, ncols, dev=None, ce_blowup=1):
        self.ctx, self.log_n, self.log_b, self.ncols = ctx, log_n, log_blowup, ncols
        self.log_ce = log_n + (ce_blowup.bit_length() - 1)
        self.prog = E.compile_program(composition(ncols, log_n, ce_blowup), ncols, lde_step=ce_blowup, log_ce=self.log_ce)

    def run(self, lde, out):
        # lde: (ncols, N) device tensor / array holding the bit-reversed LDE; the first 2^log_ce entries
        # of every column are the ce-domain evaluations in bit-reversed order (src/prover.rs:86-91)
        N = 1 << (self.log_n + self.log_b)
        self.ctx.eval_constraints(self.prog, out, self.log_ce, base_cols=lde, nbase=self.ncols, base_stride=N,
                                  fq_field=FP, offset=GENERATOR, trace_bitrev=True)
