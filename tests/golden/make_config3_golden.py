"""Generates tests/golden/config3.json: the Merkle root and the digest of the constraint-evaluation column that
bench.py's step must produce, computed by the CPU oracle (oracle/gl_oracle.c) — TEST INFRASTRUCTURE, run offline:

    python tests/golden/make_config3_golden.py            # every case below (the 2^24 one: ~10 min on 8 cores, ~20 GiB)
    python tests/golden/make_config3_golden.py --max-log-n 20

Cases: the workload of BASELINE config 3 exactly as bench.py builds it — ms_fill_random(seed 3000 + rank) per 32-column
block, iNTT, coset LDE x8 in bit-reversed row order, SHA-256 leaves and heap-layout nodes (src/prover.rs:46-55,
src/merkle.rs:412-508), the synthetic composition of ministark_b200/synth_air.py — at 2^24 x 32 (one GPU) and at the
small sharded shapes bench.py --gpus N re-checks outside its timed region (2^16 rows x 32*N columns).

The LDE is streamed coset by coset (block q of the bit-reversed LDE = the size-n NTT over offset * g_N^bitrev(q), rows
bit-reversed), so the 32 GiB matrix never exists on the host.  The streaming path is checked against the oracle's direct
orc.lde at the small sizes before its 2^24 answer is trusted (check_streaming)."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from oracle import synth_oracle  # noqa: E402

P = 2**64 - 2**32 + 1


def device_fill_random_range(i0, count, seed):
    """ms_fill_random (csrc/api_core.cu fill_random_kernel) restated, words [i0, i0 + count): word i = Montgomery form of
    the first splitmix64 draw < p of the stream seeded with seed ^ (0xD1B54A32D192ED03 * (i + 1))."""
    with np.errstate(over="ignore"):
        i = np.arange(i0 + 1, i0 + count + 1, dtype=np.uint64)
        s = np.uint64(seed) ^ (np.uint64(0xD1B54A32D192ED03) * i)
        del i
        out = np.empty(count, dtype=np.uint64)
        todo = None
        while True:
            cur = s if todo is None else s[todo]
            cur = cur + np.uint64(0x9E3779B97F4A7C15)
            z = cur.copy()
            z ^= z >> np.uint64(30)
            z *= np.uint64(0xBF58476D1CE4E5B9)
            z ^= z >> np.uint64(27)
            z *= np.uint64(0x94D049BB133111EB)
            z ^= z >> np.uint64(31)
            good = z < np.uint64(P)
            if todo is None:
                s = cur
                out[:] = z
                bad = np.nonzero(~good)[0]
            else:
                s[todo] = cur
                out[todo[good]] = z[good]
                bad = todo[~good]
            if bad.size == 0:
                break
            todo = bad
    return orc.to_mont(out)


def trace_blocks(log_n, ncols_per, world, seed=3000):
    """(world * ncols_per, n): rank r's block is ms_fill_random(ncols_per * n words, seed + r); built column by column"""
    n = 1 << log_n
    out = np.empty((world * ncols_per, n), dtype=np.uint64)
    for r in range(world):
        for c in range(ncols_per):
            out[r * ncols_per + c] = device_fill_random_range(c * n, n, seed + r)
    return out


def brev(v, bits):
    return int(format(v, f"0{bits}b")[::-1], 2) if bits else 0


def streamed_commit(polys, log_n, log_b):
    """root of the bit-reversed LDE's tree, one coset block at a time; also returns block 0 (the ce-domain prefix)"""
    n, nb = 1 << log_n, 1 << log_b
    ncols = polys.shape[0]
    gN = orc.root_of_unity(log_n + log_b)
    leaves = np.empty((n * nb, 32), dtype=np.uint8)
    block0 = None
    for q in range(nb):
        off = orc.fp_mul(orc.generator(), orc.fp_pow(gN, brev(q, log_b)))
        ev = orc.ntt(polys, 1, log_n, off)                       # natural order over the coset
        for c in range(ncols):
            orc.lib().orc_bit_reverse(orc._p(ev[c]), 1, log_n)   # rows of the block in bit-reversed order
        leaves[q * n:(q + 1) * n] = orc.hash_rows(ev, 1)
        if q == 0:
            block0 = ev
    nodes = orc.merkle_nodes(leaves)
    root = nodes[1].tobytes()
    del nodes, leaves
    return root, block0


def case(log_n, ncols_per, world, log_b=3, check_streaming=False):
    t0 = time.time()
    trace = trace_blocks(log_n, ncols_per, world)
    trace_sha = hashlib.sha256(trace.tobytes() if log_n <= 20 else memoryview(trace.reshape(-1))).hexdigest()
    polys = orc.ntt(trace, 1, log_n, inverse=True)
    del trace
    polys0_sha = hashlib.sha256(polys[0].tobytes()).hexdigest()
    root, block0 = streamed_commit(polys, log_n, log_b)
    if check_streaming:
        lde = orc.lde(polys, 1, log_n, log_b, orc.generator(), bitrev=True)
        want = orc.merkle_nodes(orc.hash_rows(lde, 1))[1].tobytes()
        assert want == root, "streamed commitment differs from the direct oracle path"
        assert np.array_equal(lde[:, :1 << log_n], block0)
        del lde
    del polys
    # composition: every rank evaluates the 32-column AIR on its own block; the partial columns are summed
    total = None
    for r in range(world):
        ce = synth_oracle.constraint_eval(orc, block0[r * ncols_per:(r + 1) * ncols_per], log_n, log_b, ncols_per)
        ce = np.ascontiguousarray(ce).reshape(-1)
        total = ce if total is None else orc.pointwise("add", total, 1, ce, 1)
    out = {
        "log_n": log_n, "ncols_per_rank": ncols_per, "world": world, "log_blowup": log_b, "seed": 3000,
        "trace_sha256": trace_sha,
        "polys_col0_sha256": polys0_sha,
        "lde_block0_col0_sha256": hashlib.sha256(block0[0].tobytes()).hexdigest(),
        "merkle_root": root.hex(),
        "constraint_eval_sha256": hashlib.sha256(np.ascontiguousarray(total).tobytes()).hexdigest(),
        "constraint_eval_first": [int(x) for x in np.ascontiguousarray(total).reshape(-1)[:4]],
    }
    print(f"case 2^{log_n} x {ncols_per}*{world}: {time.time() - t0:.1f} s, root {root.hex()[:16]}", flush=True)
    return out


CASES = [(12, 32, 1), (16, 32, 1), (16, 32, 2), (16, 32, 4), (16, 32, 8), (20, 32, 1), (24, 32, 1)]


def key(log_n, ncols_per, world):
    return f"2^{log_n}x{ncols_per}x{world}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log-n", type=int, default=24)
    ap.add_argument("--out", default=os.path.join(HERE, "config3.json"))
    args = ap.parse_args()
    res = json.load(open(args.out)) if os.path.exists(args.out) else {}
    for log_n, ncols_per, world in CASES:
        if log_n > args.max_log_n:
            continue
        res[key(log_n, ncols_per, world)] = case(log_n, ncols_per, world, check_streaming=log_n <= 16)
        json.dump(res, open(args.out, "w"), indent=1, sort_keys=True)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
