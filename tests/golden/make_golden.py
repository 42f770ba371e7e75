#!/usr/bin/env python
"""Regenerates tests/golden/golden_r01.json.

The reference ships no numeric golden vectors for this path and cannot be built here, so these are
SELF-GENERATED regression pins: SHA-256 digests of what the CPU oracle (oracle/gl_oracle.c, oracle/stark_oracle.py)
produces for seeded inputs, frozen at the state in which the oracle agreed with the big-integer spec, the O(n^2) DFT
definition, hashlib, the restated verifier and the GPU.  They guard the oracle — the checker of every parity test —
against silent drift; they are NOT evidence of parity with the reference binary.

    python tests/golden/make_golden.py          # rewrites the JSON next to this script
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from oracle import stark_oracle as SO  # noqa: E402


def h(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compute():
    from ministark_b200.air import Air, ProofOptions
    from ministark_b200.examples import brainfuck as bf
    from ministark_b200.examples import fib, perm
    g = {}
    m = orc.rand_matrix(3, 1 << 10, 1, seed=2024)
    g["ntt_fwd_subgroup_3x2p10"] = h(orc.ntt(m, 1, 10))
    g["ntt_fwd_coset_3x2p10"] = h(orc.ntt(m, 1, 10, orc.generator()))
    g["ntt_inv_coset_3x2p10"] = h(orc.ntt(m, 1, 10, orc.generator(), inverse=True))
    q = orc.rand_matrix(2, 1 << 8, 3, seed=7)
    g["ntt_fwd_fq3_2x2p8"] = h(orc.ntt(q, 3, 8))
    polys = orc.ntt(m, 1, 10, inverse=True)
    lde = orc.lde(polys, 1, 10, 3, orc.generator(), bitrev=True)
    g["lde_bitrev_3x2p10_x8"] = h(lde)
    g["merkle_root_3x2p13"] = orc.merkle_nodes(orc.hash_rows(lde, 1))[1].tobytes().hex()
    cw = orc.rand_matrix(1, 1 << 12, 3, seed=11)[0]
    alpha = orc.rand_matrix(1, 1, 3, seed=12)[0]
    g["fri_fold_fq3_2p12_ff8"] = h(orc.fri_apply_drp(cw, 3, 12, 3, alpha))
    g["scan_affine_fq3_1000"] = h(orc.scan_affine(3, 1000, alpha, a=orc.rand_matrix(1, 1000, 3, seed=13)[0], fa=3,
                                                   b=orc.rand_matrix(1, 1000, 1, seed=14)[0], fb=1, inclusive=True))
    g["pow_nonce_12bits"] = orc.pow_grind(hashlib.sha256(b"golden").digest(), 12)
    coin = SO.Coin(hashlib.sha256(b"golden coin").digest(), 3)
    g["coin_draws_fq3"] = [list(coin.draw()) for _ in range(3)]
    g["coin_queries"] = coin.draw_queries(8, 1 << 20)
    trace, last = fib.gen_trace(8 << 7)
    claim = fib.FibClaim(last)
    mk = lambda n, o: Air(claim.AirConfig, n, claim.get_public_inputs(), ProofOptions(*o))
    proof = SO.cpu_prove(claim, (32, 4, 8, 8, 64), trace.base_columns(), mk)
    g["fib_2p7_rows_claim"] = last
    g["fib_2p7_rows_proof_sha256"] = hashlib.sha256(proof).hexdigest()
    g["fib_2p7_rows_proof_len"] = len(proof)
    pclaim = perm.PermClaim()
    ptr = perm.gen_trace(1 << 6)
    pmk = lambda n, o: Air(pclaim.AirConfig, n, pclaim.get_public_inputs(), ProofOptions(*o))
    pproof = SO.cpu_prove(pclaim, (12, 8, 4, 4, 8), ptr.base_columns(), pmk, ext_builder=ptr.build_extension_columns)
    g["perm_2p6_rows_proof_sha256"] = hashlib.sha256(pproof).hexdigest()
    btrace, out = bf.simulate(bf.HELLO_WORLD)
    bclaim = bf.BrainfuckClaim(bf.HELLO_WORLD, b"", out)
    bmk = lambda n, o: Air(bclaim.AirConfig, n, bclaim, ProofOptions(*o))
    bproof = SO.cpu_prove(bclaim, (19, 16, 20, 16, 16), btrace.base_columns(), bmk, ext_builder=btrace.build_extension_columns)
    g["brainfuck_hello_world_base_trace_sha256"] = h(btrace.base_columns())
    g["brainfuck_hello_world_proof_sha256"] = hashlib.sha256(bproof).hexdigest()
    g["brainfuck_hello_world_proof_len"] = len(bproof)
    return g


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_r01.json")
    json.dump(compute(), open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out)
