"""CPU-only: the host-side index logic of the multi-GPU prover (ministark_b200/prover_mgpu.py) — which digests a batched
Merkle proof names (the walk of src/merkle.rs:149-207), who owns each of them when the tree is split into G subtrees, and the
top levels rebuilt from the subtree roots — against the oracle's single-tree MerkleTreeImpl::prove restatement."""
import hashlib
import random

import numpy as np
import pytest

from ministark_b200 import prover_mgpu as M


@pytest.mark.parametrize("n,G", [(64, 2), (64, 4), (256, 8), (8, 4), (16, 2)])
def test_sharded_merkle_view_equals_single_tree(orc, n, G):
    from oracle import stark_oracle as SO
    rng = np.random.default_rng(n + G)
    leaves = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    nodes = orc.merkle_nodes(leaves)
    log_g = G.bit_length() - 1
    per = n // G
    # every rank's own subtree (heap layout, root at local index 1) from its rows alone
    local_nodes = [orc.merkle_nodes(leaves[r * per:(r + 1) * per]) for r in range(G)]
    top = M.top_levels([ln[1].tobytes() for ln in local_nodes])
    assert top[1] == nodes[1].tobytes()
    for k in range(1, G):
        assert top[k] == nodes[k].tobytes()
    r = random.Random(n)
    for trial in range(20):
        ids = sorted(set(r.randrange(n) for _ in range(r.randrange(1, 12))))
        init, sib, path = M.merkle_walk(n, ids)
        want = SO._merkle_prove(leaves, nodes, ids)
        got_init = [leaves[i].tobytes() for i in init]
        got_sib = [leaves[i].tobytes() for i in sib]
        got_path = []
        for k in path:
            owner, loc = M.node_owner(k, log_g)
            # (heap index 0 is the unused default digest the reference's walk names when it passes node 3, src/merkle.rs:441)
            got_path.append((top[loc] if loc else bytes(32)) if owner is None else local_nodes[owner][loc].tobytes())
        assert (got_path, got_init, got_sib) == (want["nodes"], want["initial_leaves"], want["sibling_leaves"])
        # leaf ownership used by the prover: global leaf i lives on rank i // per at local index i % per
        assert all(leaves[i].tobytes() == leaves[(i // per) * per + i % per].tobytes() for i in init + sib)


def test_block_offsets_match_the_lde_plan():
    """block q of a bit-reversed LDE is the coset offset * g_N^bitrev(q) (csrc/api_ntt.cu, src/matrix.rs:225-234)"""
    P = M.P
    for log_n, log_b in [(4, 2), (6, 3), (10, 4)]:
        gN = M.domain_generator(log_n + log_b)
        gn = M.domain_generator(log_n)
        n, beta = 1 << log_n, 1 << log_b
        # position p of the bit-reversed LDE holds the evaluation at offset * g_N^bitrev_{log_N}(p)
        for p in (0, 1, n - 1, n, 3 * n + 5, beta * n - 1):
            q, t = divmod(p, n)
            e = M._brev(p, log_n + log_b)
            x = 7 * pow(gN, e, P) % P
            h = 7 * pow(gN, M._brev(q, log_b), P) % P
            assert x == h * pow(gn, M._brev(t, log_n), P) % P
