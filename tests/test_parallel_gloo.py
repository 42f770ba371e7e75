"""CPU, world_size 2 over gloo: the multi-GPU host logic of ministark_b200/parallel.py (column
blocks -> all-to-all row slabs -> subtree roots -> all-gather -> merged root) with the CPU oracle
plugged in as the compute engine.  The merged root must equal the single-process root of the
full matrix — the sharded commitment is bit-identical to the reference's single tree."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """CPU stand-in for CudaEngine: torch CPU tensors + oracle/gl_oracle.c"""

    def __init__(self):
        import torch
        from oracle import oracle
        self.torch, self.orc = torch, oracle

    def empty(self, shape):
        return self.torch.zeros(shape, dtype=self.torch.int64)

    def view(self, buf, col, lo, hi):
        return buf[col, lo:hi]

    def _np(self, t):
        return t.numpy().view(np.uint64)

    def intt(self, src, dst, log_n, ncols):
        dst.copy_(self.torch.from_numpy(self.orc.ntt(self._np(src), 1, log_n, inverse=True).view(np.int64)))

    def lde(self, coeffs, out, log_n, log_b, ncols):
        out.copy_(self.torch.from_numpy(self.orc.lde(self._np(coeffs), 1, log_n, log_b, self.orc.generator(), True).view(np.int64)))

    def subtree_root(self, slab, nrows, ncols):
        leaves = self.orc.hash_rows(np.ascontiguousarray(self._np(slab)), 1)
        if nrows == 1:
            return leaves[0].tobytes()
        return self.orc.merkle_nodes(leaves)[1].tobytes()

    def gather_digests(self, dist, digest, world):
        t = self.torch.frombuffer(bytearray(digest), dtype=self.torch.uint8)
        out = [self.torch.empty(32, dtype=self.torch.uint8) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.numpy().tobytes() for o in out]


class PeerOracleEngine(OracleEngine):
    """OracleEngine + the fused-exchange hooks, with "peer memory" played by POSIX shared memory: ipc_export hands out the
    segment name, ipc_open attaches to a peer's segment, and addresses are (segment index << 48 | byte offset) so that the
    block-pointer arithmetic of ShardedCommit.lde_columns is exercised exactly as with device addresses"""
    peer_scatter = True

    def __init__(self):
        super().__init__()
        self.segments = []          # index -> (SharedMemory, numpy uint64 view)

    def _register(self, shm):
        self.segments.append((shm, np.ndarray(shm.size // 8, dtype=np.uint64, buffer=shm.buf)))
        return len(self.segments) << 48

    def alloc_exportable(self, shape):
        from multiprocessing import shared_memory
        shm = shared_memory.SharedMemory(create=True, size=int(np.prod(shape)) * 8)
        base = self._register(shm)
        arr = np.ndarray(shape, dtype=np.int64, buffer=shm.buf)
        arr[:] = 0
        self._own = shm
        return self.torch.from_numpy(arr), base

    def ipc_export(self, ptr):
        return self._own.name

    def ipc_open(self, handle):
        from multiprocessing import shared_memory
        return self._register(shared_memory.SharedMemory(name=handle))

    LOCAL = 1 << 60                 # "address" space of the local LDE buffer

    def ptr(self, buf):
        self._lde = buf
        return self.LOCAL

    def sync(self):
        pass

    def rendezvous(self, dist):
        dist.barrier()              # the CPU engine is synchronous: a host barrier is the whole rendezvous

    def ipc_close(self, ptr):
        self.closed = getattr(self, "closed", 0) + 1

    def free_exportable(self, ptr):
        self.freed = True

    def _store(self, addr, words):
        local = bool(addr & self.LOCAL)
        addr &= self.LOCAL - 1
        seg, off = addr >> 48, (addr & ((1 << 48) - 1)) // 8
        target = self._np(self._lde).reshape(-1) if local else self.segments[seg - 1][1]
        target[off:off + words.size] = words

    def lde_scatter(self, coeffs, work_ptr, log_n, log_b, ncols, block_ptrs, block_stride, dup_ptrs, dup_stride):
        n = 1 << log_n
        lde = self.orc.lde(self._np(coeffs), 1, log_n, log_b, self.orc.generator(), True)
        for q, (bp, dp) in enumerate(zip(block_ptrs, dup_ptrs)):
            for c in range(ncols):
                blk = lde[c, q * n:(q + 1) * n]
                self._store(bp + c * block_stride * 8, blk)
                if dp:
                    self._store(dp + c * dup_stride * 8, blk)

    def close(self):
        for shm, arr in self.segments:
            del arr
        self.segments = []
        try:
            self._own.close()
            self._own.unlink()
        except Exception:
            pass


def _worker(rank, world, port, log_n, log_b, ncols, q, fused=False):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from ministark_b200 import parallel
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = orc.rand_matrix(ncols, 1 << log_n, 1, seed=77)          # every rank can regenerate the trace
        eng = PeerOracleEngine() if fused else OracleEngine()
        sc = parallel.ShardedCommit(eng, dist, log_n, log_b, ncols, fused=True if fused else None)
        assert sc.fused == bool(fused)
        lo, hi = sc.lo, sc.hi
        import torch
        local = torch.from_numpy(full[lo:hi].view(np.int64).copy())
        sc.transform(local)
        root = sc.commit()
        if fused:
            sc.transform(local)                      # second round: barrier before the slabs are overwritten
            assert sc.commit() == root
            n = 1 << log_n                           # the local copy of block 0 (ce-domain prefix)
            want = orc.lde(orc.ntt(full[lo:hi], 1, log_n, inverse=True), 1, log_n, log_b, orc.generator(), True)
            assert np.array_equal(sc.lde.numpy().view(np.uint64)[:, :n], want[:, :n])
            sc.close()                               # unmaps the peers' slabs, frees its own, falls back to unfused state
            assert not sc.fused and eng.closed == world - 1 and eng.freed
            eng.close()
        q.put((rank, root))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("log_n,log_b,ncols,fused", [(6, 3, 8, False), (5, 1, 2, False), (6, 3, 8, True), (5, 1, 4, True)])
def test_sharded_commit_world2_matches_single_tree(orc, log_n, log_b, ncols, fused):
    """fused=True: the exchange-free path (LDE blocks stored straight into the owners' slabs) with shared memory
    standing in for CUDA IPC peer memory"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, log_n, log_b, ncols, q, fused)) for r in range(2)]
    for p in procs:
        p.start()
    roots = dict(q.get(timeout=60) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = orc.rand_matrix(ncols, 1 << log_n, 1, seed=77)
    lde = orc.lde(orc.ntt(full, 1, log_n, inverse=True), 1, log_n, log_b, orc.generator(), True)
    want = orc.merkle_nodes(orc.hash_rows(lde, 1))[1].tobytes()
    assert roots[0] == roots[1] == want


def test_merge_subtree_roots_and_blocks(orc):
    from ministark_b200 import parallel
    leaves = orc.hash_rows(orc.rand_matrix(3, 16, 1, seed=1), 1)
    nodes = orc.merkle_nodes(leaves)
    # subtree roots of 4 slabs of 4 leaves are the level-2 nodes 4..7 of the heap layout
    subs = [orc.merkle_nodes(leaves[4 * d:4 * d + 4])[1].tobytes() for d in range(4)]
    assert subs == [nodes[4 + d].tobytes() for d in range(4)]
    assert parallel.merge_subtree_roots(subs) == nodes[1].tobytes()
    assert parallel.merge_subtree_roots(subs[:1]) == subs[0]
    assert [parallel.column_block(32, 4, r) for r in range(4)] == [(0, 8), (8, 16), (16, 24), (24, 32)]
    with pytest.raises(AssertionError):
        parallel.merge_subtree_roots(subs[:3])
