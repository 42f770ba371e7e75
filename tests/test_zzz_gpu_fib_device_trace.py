"""GPU: examples/fib with its trace built in device memory (fib.gen_trace(..., device=...): the sequence v_k = 2^F(k) has
period 96 over Goldilocks because 2 has order 192, so the (8, n/8) matrix is a 12-row block tiled on the device).  The
resident tensor must equal the host construction word for word and yield the same proof bytes."""
import numpy as np
import pytest

from ministark_b200.examples import fib
from ministark_b200.prover import GpuProver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_rows", [7, 10, 16])
def test_device_trace_equals_host_trace_and_proves_the_same(log_rows):
    torch = pytest.importorskip("torch")
    host, last = fib.gen_trace(8 << log_rows)
    dev, last_dev = fib.gen_trace(8 << log_rows, device=torch.device("cuda", 0))
    cols = dev.base_columns()
    assert cols.is_cuda and cols.is_contiguous() and tuple(cols.shape) == (8, 1 << log_rows) and last_dev == last
    assert np.array_equal(cols.cpu().numpy().view(np.uint64), host.base_columns())
    claim = fib.FibClaim(last)
    prover = GpuProver.shared(0)
    assert prover.prove(claim, fib.OPTIONS, dev).to_bytes() == prover.prove(claim, fib.OPTIONS, host).to_bytes()
