"""synth_oracle.py — CPU-oracle side of the synthetic config-3 AIR.  TEST INFRASTRUCTURE ONLY.
States the same AIR as ministark_b200/synth_air.py independently, in the oracle's tuple format
(composition per src/air.rs:50-82)."""
import numpy as np

from . import eval_oracle
from . import pyspec as S

P = S.P


def _coeff(k, which):
    return (0x9E3779B97F4A7C15 * (2 * k + which + 1)) % P


def composition_tree(ncols, log_n, ce_blowup=1):
    n = 1 << log_n
    g_inv = pow(S.root_of_unity(log_n), -1, P)
    x = ('x',)
    const = lambda v: ('const', (v % P, 0, 0), False)
    sub = lambda a, b: ('add', a, ('neg', b))
    num = sub(x, const(g_inv))
    den = sub(('pow', x, n), const(1))
    adj = (n * ce_blowup - 1) - ((2 * (n - 1) + 1) - n)
    xadj = ('pow', x, adj)
    total = None
    for k in range(ncols):
        c = sub(('trace', k, 1), ('mul', ('trace', (k - 2) % ncols, 0), ('trace', (k - 1) % ncols, 0)))
        c = ('div', ('mul', c, num), den)
        term = ('mul', c, ('add', ('mul', xadj, const(_coeff(k, 0))), const(_coeff(k, 1))))
        total = term if total is None else ('add', total, term)
    return total


def constraint_eval(orc, lde, log_n, log_blowup, ncols, ce_blowup=1):
    """CPU path: bit_reverse_ce_trace (src/prover.rs:185-194) then eval_cpu::eval."""
    log_ce = log_n + (ce_blowup.bit_length() - 1)
    m = 1 << log_ce
    cols = np.stack([orc.bit_reverse(np.ascontiguousarray(lde[c][:m]), 1, log_ce) for c in range(ncols)])
    return eval_oracle.evaluate(composition_tree(ncols, log_n, ce_blowup), log_ce, orc.generator(), cols,
                                lde_step=ce_blowup)
