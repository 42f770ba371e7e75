"""GPU side of the golden fixtures (tests/golden/golden_r01.json, see tests/test_golden.py): the device must reproduce
the digests the CPU oracle and the CPU prover froze.  (Named to run after the per-kernel parity tests.)"""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden_r01.json")))


@pytest.mark.gpu
def test_gpu_reproduces_the_golden_fixtures(orc):
    import ministark_b200 as ms
    from ministark_b200.air import ProofOptions
    from ministark_b200.examples import brainfuck as bf
    from ministark_b200.examples import fib
    from ministark_b200.prover import GpuProver
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    ctx = ms.Context(0)
    m = orc.rand_matrix(3, 1 << 10, 1, seed=2024)
    a = m.copy()
    ctx.ntt_batch(a, ms.FP, 10, 3, offset=ms.GENERATOR)
    assert h(a) == GOLDEN["ntt_fwd_coset_3x2p10"]
    polys = m.copy()
    ctx.ntt_batch(polys, ms.FP, 10, 3, inverse=True)
    lde = np.empty((3, 1 << 13), dtype=np.uint64)
    ctx.lde_batch(polys, lde, ms.FP, 10, 3, 3, offset=ms.GENERATOR, bitrev=True)
    assert h(lde) == GOLDEN["lde_bitrev_3x2p10_x8"]
    assert ctx.merkle_commit(lde, ms.FP, 1 << 13, 3).hex() == GOLDEN["merkle_root_3x2p13"]
    assert ctx.pow_grind(hashlib.sha256(b"golden").digest(), 12) == GOLDEN["pow_nonce_12bits"]
    prover = GpuProver.shared(0)
    trace, last = fib.gen_trace(8 << 7)
    assert last == GOLDEN["fib_2p7_rows_claim"]
    proof = prover.prove(fib.FibClaim(last), ProofOptions(32, 4, 8, 8, 64), trace).to_bytes()
    assert hashlib.sha256(proof).hexdigest() == GOLDEN["fib_2p7_rows_proof_sha256"]
    btrace, out = bf.simulate(bf.HELLO_WORLD)
    assert h(btrace.base_columns()) == GOLDEN["brainfuck_hello_world_base_trace_sha256"]
    bproof = prover.prove(bf.BrainfuckClaim(bf.HELLO_WORLD, b"", out), bf.OPTIONS, btrace).to_bytes()
    assert hashlib.sha256(bproof).hexdigest() == GOLDEN["brainfuck_hello_world_proof_sha256"]
