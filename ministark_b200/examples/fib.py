"""examples/fib (examples/fib/main.rs): "multiplicative Fibonacci" over Goldilocks, 8 columns, Fq = Fp.

    FibAirConfig.constraints   examples/fib/main.rs:78-150   (8 boundary + 1 terminal + 8 transition constraints)
    gen_trace                  examples/fib/main.rs:175-222
    OPTIONS                    examples/fib/main.rs:225        ProofOptions(32, 4, 8, 8, 64)

Values are canonical integers; the trace matrix is returned as Montgomery words (the reference's in-memory form).
"""
import numpy as np

from .. import expr as E
from ..air import AirConfig, ProofOptions, domain_generator
from ..prover import Stark, Trace

P = E.P
OPTIONS = ProofOptions(32, 4, 8, 8, 64)
SECURITY_LEVEL = 30


class FibAirConfig(AirConfig):
    NUM_BASE_COLUMNS = 8
    FQ_IS_FP = True

    @staticmethod
    def gen_hints(trace_len, claimed_nth_fib_number, challenges):
        return [claimed_nth_fib_number]                     # FibHint::ClaimedNthFibNum = 0

    @staticmethod
    def constraints(trace_len):
        log_n = trace_len.bit_length() - 1
        g = domain_generator(log_n)
        x = E.X()
        first_trace_x = E.Constant(1)                       # trace_xs.element(0)
        last_trace_x = E.Constant(pow(g, trace_len - 1, P))  # trace_xs.element(n - 1) = g^-1
        one = E.Constant(1)
        v = [one, one + one]
        v.append(v[1] * v[0])
        for i in range(3, 8):
            v.append(v[i - 2] * v[i - 1])
        T = E.Trace
        boundary = [(T(i, 0) - v[i]) / (x - first_trace_x) for i in range(8)]
        terminal = [(T(7, 0) - E.Hint(0)) / (x - last_trace_x)]
        transition_exprs = [
            T(0, 1) - T(6, 0) * T(7, 0),
            T(1, 1) - T(7, 0) * T(0, 1),
            T(2, 1) - T(0, 1) * T(1, 1),
            T(3, 1) - T(1, 1) * T(2, 1),
            T(4, 1) - T(2, 1) * T(3, 1),
            T(5, 1) - T(3, 1) * T(4, 1),
            T(6, 1) - T(4, 1) * T(5, 1),
            T(7, 1) - T(5, 1) * T(6, 1),
        ]
        # all rows except the last: multiply by (x - t_(n-1)) / (x^n - 1)
        transition = [c * ((x - last_trace_x) / (x ** trace_len - one)) for c in transition_exprs]
        return boundary + terminal + transition


def gen_trace(n, pinned=False, device=None):
    """n = total number of sequence values; the trace has n / 8 rows of 8 consecutive values (main.rs:175-222).
    v_0 = 1, v_1 = 2, v_k = v_(k-2) * v_(k-1).  Returns (Trace, last value of column 7 as a canonical int).
    pinned: put the columns in page-locked host memory (the reference allocates them with GpuAllocator, main.rs:181-188)
    so that the prover's chunked upload overlaps with the transforms.
    device: build the columns ON that device (SURVEY.md §8f rank 3 for this example): only the period of the sequence —
    a few rows, see below — is computed on the host and uploaded, the (8, n/8) matrix is tiled from it in device memory
    and handed to the prover as a resident tensor; nothing of size n crosses PCIe."""
    assert n & (n - 1) == 0 and n > 8
    num_rows = n // 8
    # v_k = 2^F(k) and 2 has order 192 in Goldilocks, so the sequence is periodic (period 96 = Pisano(192));
    # generate rows by the recurrence until the first row repeats, then tile — identical values, O(period) big-int work
    rows, v = [], [1, 2]
    for i in range(2, 8):
        v.append(v[i - 2] * v[i - 1] % P)
    while len(rows) < num_rows:
        if rows and v == rows[0]:
            break
        rows.append(list(v))
        nv = [v[6] * v[7] % P]
        nv.append(v[7] * nv[0] % P)
        for i in range(2, 8):
            nv.append(nv[i - 2] * nv[i - 1] % P)
        v = nv
    period = np.array([[x * 2**64 % P for x in r] for r in rows], dtype=np.uint64)          # Montgomery words
    reps = -(-num_rows // len(rows))
    last = rows[(num_rows - 1) % len(rows)][7]
    if device is not None:
        import torch
        block = torch.from_numpy(period.view(np.int64)).to(device)                             # (period rows, 8)
        cols = block.repeat(reps, 1)[:num_rows].t().contiguous()
        if cols.is_cuda:        # the prover reads the tensor on its own stream: hand it over complete
            torch.cuda.current_stream(cols.device).synchronize()
        return Trace(cols), last
    cols = np.ascontiguousarray(np.tile(period, (reps, 1))[:num_rows].T)
    if pinned:
        import torch
        host = torch.empty((8, num_rows), dtype=torch.int64, pin_memory=True)
        host.numpy().view(np.uint64)[:] = cols
        return Trace(host), last
    return Trace(cols), last


class FibClaim(Stark):
    AirConfig = FibAirConfig

    def __init__(self, claimed_value):
        self.claimed_value = int(claimed_value)

    def get_public_inputs(self):
        return self.claimed_value
