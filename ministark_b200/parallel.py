"""parallel.py — one-process-per-GPU sharding of the commitment pipeline.

The reference has no multi-device code at all (one Metal device, gpu/src/plan.rs:465-469).  The
path shards like this:

  phase A  (column sharded, no communication)   rank r owns a block of columns:
           iNTT + coset LDE of its columns                                 (src/prover.rs:46-51)
           partial composition: the constraints that read its columns      (src/air.rs:50-82 is a SUM
           over constraints, so it splits by constraint; column-local constraints need no halo)
  exchange all-to-all: rank r sends rows [d*N/G, (d+1)*N/G) of its columns to rank d — a leaf hash
           needs every column of its row (src/hash.rs:92-99), so some row-wise exchange is
           unavoidable; volume (G-1)/G of the LDE per GPU over NVLink
  phase B  (row sharded)  leaf hashes + Merkle subtree of the local row slab   (src/merkle.rs:412-508)
  commit   all-gather of the G subtree roots (32 B each); every rank finishes the top log2(G)
           levels -> the same root as the single-device tree               (src/merkle.rs:485-508)
           all-gather of the partial composition columns, summed locally.

The compute is delegated to an `engine` (the CUDA Context in production; the CPU tests plug a
numpy engine in to exercise exactly this host logic under gloo with world_size 2).
Collectives go through torch.distributed (NCCL over NVLink on the GPU box).
"""
import hashlib

import numpy as np


def column_block(ncols_total, world, rank):
    """block ownership: rank r owns global columns [lo, hi); blocks keep the global column order in the
    row slabs without a permutation"""
    per = (ncols_total + world - 1) // world
    lo = min(rank * per, ncols_total)
    return lo, min(lo + per, ncols_total)


def merge_subtree_roots(roots):
    """top log2(G) levels of the heap-layout tree from the G subtree roots (left to right):
    nodes[k] = SHA-256(nodes[2k] || nodes[2k+1])  (src/merkle.rs:499-506, src/hash.rs:77-82)"""
    level = [bytes(r) for r in roots]
    g = len(level)
    assert g >= 1 and g & (g - 1) == 0, "world size must be a power of two"
    while len(level) > 1:
        level = [hashlib.sha256(level[2 * i] + level[2 * i + 1]).digest() for i in range(len(level) // 2)]
    return level[0]


class ShardedCommit:
    """Column-sharded LDE -> all-to-all -> row-sharded Merkle commit.

    engine must provide (device arrays are whatever `engine.empty` returns):
        empty(shape_words) -> buffer              uint64 words, resident where the engine computes
        view(buf, col, row_lo, row_hi) -> buffer  contiguous view of rows [row_lo,row_hi) of column col
                                                  of a (ncols, nrows) column-major buffer
        intt(src, dst, log_n, ncols)
        lde(coeffs, out, log_n, log_blowup, ncols)
        subtree_root(slab, nrows, ncols) -> 32 bytes       leaves+nodes of the slab, root = nodes[1];
                                                           for nrows == 1 the single leaf digest
    """

    def __init__(self, engine, dist, log_n, log_blowup, ncols_total, polys=None, lde=None):
        self.e, self.dist = engine, dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        if self.world & (self.world - 1):
            raise ValueError("world size must be a power of two")
        self.log_n, self.log_b, self.ncols_total = log_n, log_blowup, ncols_total
        self.lo, self.hi = column_block(ncols_total, self.world, self.rank)
        self.nloc = self.hi - self.lo
        self.N = 1 << (log_n + log_blowup)
        if self.N % self.world or ncols_total % self.world:
            raise ValueError("rows and columns must divide evenly over the ranks")
        self.rows_per = self.N // self.world
        n = 1 << log_n
        self.polys = polys if polys is not None else engine.empty((self.nloc, n))
        self.lde = lde if lde is not None else engine.empty((self.nloc, self.N))
        self.slab = engine.empty((ncols_total, self.rows_per)) if self.world > 1 else None

    def transform(self, trace):
        """phase A on the local columns: trace (nloc, n) -> self.polys, self.lde"""
        self.e.intt(trace, self.polys, self.log_n, self.nloc)
        self.e.lde(self.polys, self.lde, self.log_n, self.log_b, self.nloc)
        return self.lde

    def exchange(self):
        """all-to-all into row slabs; returns the (ncols_total, rows_per) slab of this rank"""
        if self.world == 1:
            return self.lde
        per = self.nloc
        sends, recvs = [], []
        for c in range(per):
            for d in range(self.world):
                sends.append(self.e.view(self.lde, c, d * self.rows_per, (d + 1) * self.rows_per))
        for c in range(per):
            for s in range(self.world):
                # rows of global column s*per + c arrive from rank s
                recvs.append(self.e.view(self.slab, s * per + c, 0, self.rows_per))
        # the all-to-all as one batch of point-to-point transfers (NCCL groups them into a single
        # collective-like launch over NVLink; gloo, used by the CPU tests, has no alltoall)
        ops = []
        for c in range(per):
            for d in range(self.world):
                snd, rcv = sends[c * self.world + d], recvs[c * self.world + d]
                if d == self.rank:
                    rcv.copy_(snd)
                else:
                    ops.append(self.dist.P2POp(self.dist.isend, snd, d))
                    ops.append(self.dist.P2POp(self.dist.irecv, rcv, d))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
        return self.slab

    def commit(self):
        """phase B + root merge; returns the 32-byte Merkle root of the full (ncols_total x N) LDE"""
        slab = self.exchange()
        sub = self.e.subtree_root(slab, self.rows_per, self.ncols_total)
        if self.world == 1:
            return sub
        roots = self.e.gather_digests(self.dist, sub, self.world)
        return merge_subtree_roots(roots)


class CudaEngine:
    """engine over ministark_b200.Context with torch CUDA tensors (int64 storage of the u64 words)."""

    def __init__(self, ctx, device):
        import torch
        self.torch, self.ctx, self.device = torch, ctx, device
        self._leaves = self._nodes = None

    def empty(self, shape):
        return self.torch.empty(shape, dtype=self.torch.int64, device=self.device)

    def view(self, buf, col, lo, hi):
        return buf[col, lo:hi]

    def intt(self, src, dst, log_n, ncols):
        self.ctx.ntt_batch_to(src, dst, 1, log_n, ncols, inverse=True)

    def lde(self, coeffs, out, log_n, log_b, ncols):
        from . import GENERATOR
        self.ctx.lde_batch(coeffs, out, 1, log_n, log_b, ncols, offset=GENERATOR, bitrev=True)

    def subtree_root(self, slab, nrows, ncols):
        if self._leaves is None or self._leaves.shape[0] != nrows:
            self._leaves = self.torch.empty((nrows, 4), dtype=self.torch.int64, device=self.device)
            self._nodes = self.torch.empty((nrows, 4), dtype=self.torch.int64, device=self.device)
        if nrows == 1:
            self.ctx.hash_rows(slab, self._leaves, 1, 1, ncols, col_stride=slab.shape[1])
            self.ctx.sync()
            return self._leaves.cpu().numpy().tobytes()
        return self.ctx.merkle_commit(slab, 1, nrows, ncols, col_stride=slab.shape[1], leaves=self._leaves,
                                      nodes=self._nodes)

    def gather_digests(self, dist, digest, world):
        t = self.torch.frombuffer(bytearray(digest), dtype=self.torch.uint8).to(self.device)
        out = [self.torch.empty(32, dtype=self.torch.uint8, device=self.device) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.cpu().numpy().tobytes() for o in out]
