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
  fused    (CUDA engine, G | blow-up) the exchange disappears into the LDE: the bit-reversed LDE is 2^log_blowup coset
           blocks of n rows and a row slab is a run of whole blocks, so the last NTT pass stores every block straight
           into the slab of the GPU that hashes it — peer memory mapped with CUDA IPC, written over NVLink/NVSwitch
           (ms_lde_batch_scatter) — plus a local copy of the ce-domain prefix; two barriers per step replace the
           all-to-all (writers done before the slab is hashed; hashing done before the slab is overwritten)
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

    def __init__(self, engine, dist, log_n, log_blowup, ncols_total, polys=None, lde=None, fused=None, dup_blocks=1):
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
        self.fused, self.dup_blocks = False, dup_blocks
        nblocks = 1 << log_blowup
        can_fuse = (self.world > 1 and getattr(engine, "peer_scatter", False) and nblocks % self.world == 0 and log_n >= 4)
        if fused and not can_fuse:
            raise ValueError("fused exchange needs the CUDA engine and a world size dividing the blow-up factor")
        if can_fuse and fused is not False:
            # every rank exports its slab; block q of the bit-reversed LDE lives in the slab of rank q // (blocks per rank).
            # All ranks must end up in the same mode: a rank whose IPC setup fails makes everybody fall back to NCCL.
            ok, err = 1, None
            try:
                self.slab, self._slab_ptr = engine.alloc_exportable((ncols_total, self.rows_per))
                handle = engine.ipc_export(self._slab_ptr)
            except Exception as exc:          # noqa: BLE001 — any failure here only selects the fallback path
                ok, err, handle = 0, exc, None
            handles = [None] * self.world
            dist.all_gather_object(handles, handle)
            if ok and all(h is not None for h in handles):
                try:
                    self._peer = [self._slab_ptr if r == self.rank else engine.ipc_open(handles[r]) for r in range(self.world)]
                except Exception as exc:      # noqa: BLE001
                    ok, err = 0, exc
            else:
                ok = 0
            flags = [None] * self.world
            dist.all_gather_object(flags, ok)
            if all(flags):
                self._need_barrier = False
                self.fused = True
            elif fused:
                raise RuntimeError(f"fused exchange requested but the peer-slab setup failed on some rank: {err}")
        if not self.fused:
            self.slab = engine.empty((ncols_total, self.rows_per)) if self.world > 1 else None

    def close(self):
        """unmap the peers' slabs and free this rank's (fused mode); the object must not be used afterwards"""
        if self.fused:
            self.e.sync()
            self.dist.barrier()                      # nobody is still writing into a slab that is about to go away
            for r, p in enumerate(self._peer):
                if r != self.rank:
                    self.e.ipc_close(p)
            self.slab = None
            self.e.free_exportable(self._slab_ptr)
            self._peer, self.fused = [], False

    def transform(self, trace):
        """phase A on the local columns: trace (nloc, n) -> self.polys, self.lde"""
        self.e.intt(trace, self.polys, self.log_n, self.nloc)
        self.lde_columns(0, self.nloc)
        return self.lde

    def lde_columns(self, c0, k):
        """coset LDE of local columns [c0, c0 + k) of self.polys.  Fused mode: the blocks go straight into the owners'
        slabs (and the first `dup_blocks` blocks also into self.lde, where the local constraint evaluation reads the
        ce-domain prefix); self.lde doubles as the work buffer of the earlier passes."""
        if not self.fused:
            self.e.lde(self.polys[c0:c0 + k], self.lde[c0:c0 + k], self.log_n, self.log_b, k)
            return
        if self._need_barrier:           # nobody may still be hashing the slab this LDE is about to overwrite
            self.e.rendezvous(self.dist)
            self._need_barrier = False
        n, nblocks = 1 << self.log_n, 1 << self.log_b
        per_rank = nblocks // self.world
        col = self.lo + c0
        blocks = [self._peer[q // per_rank] + (col * self.rows_per + (q % per_rank) * n) * 8 for q in range(nblocks)]
        lde_ptr = self.e.ptr(self.lde) + c0 * self.N * 8
        dups = [lde_ptr + q * n * 8 if q < self.dup_blocks else 0 for q in range(nblocks)]
        self.e.lde_scatter(self.polys[c0:c0 + k], lde_ptr, self.log_n, self.log_b, k, blocks, self.rows_per, dups, self.N)

    def exchange(self):
        """all-to-all into row slabs; returns the (ncols_total, rows_per) slab of this rank"""
        if self.world == 1:
            return self.lde
        if self.fused:
            # this rank's stores into the peers' slabs are complete and so are everybody else's into ours: a
            # stream-ordered rendezvous (no host synchronisation)
            self.e.rendezvous(self.dist)
            self._need_barrier = True
            return self.slab
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

    def __init__(self, ctx, device, stream=None):
        """stream: the torch stream the Context launches on (its handle was given to ms.Context / set_stream); the
        collectives are ordered against torch's CURRENT stream, so the two must be the same for stream-ordered
        rendezvous — otherwise the engine falls back to host synchronisation."""
        import torch
        self.torch, self.ctx, self.device = torch, ctx, device
        self._stream_handle = None if stream is None else stream.cuda_stream
        self._leaves = self._nodes = None

    def empty(self, shape):
        return self.torch.empty(shape, dtype=self.torch.int64, device=self.device)

    def view(self, buf, col, lo, hi):
        return buf[col, lo:hi]

    # ---- peer slabs (fused exchange)
    peer_scatter = True

    def ptr(self, buf):
        return buf.data_ptr()

    def sync(self):
        self.ctx.sync()

    def rendezvous(self, dist):
        """every rank's work enqueued so far is finished before any rank's later work starts — as a one-word NCCL
        all-reduce ON THE COMPUTE STREAM (kernels before it have completed, incl. their stores into peer memory;
        kernels after it wait for it), instead of a host barrier that would drain the stream"""
        if getattr(self, "_flag", None) is None:
            self._flag = self.torch.zeros(1, dtype=self.torch.int32, device=self.device)
        if self.torch.cuda.current_stream(self.device).cuda_stream != self.stream_handle():
            self.ctx.sync()              # the context runs on a foreign stream: fall back to a full synchronisation
            dist.barrier()
            return
        dist.all_reduce(self._flag)

    def stream_handle(self):
        return getattr(self, "_stream_handle", None)

    def alloc_exportable(self, shape):
        """a cudaMalloc'ed buffer (exportable with CUDA IPC, unlike a sub-allocation of torch's caching allocator)
        wrapped as a torch tensor"""
        nwords = int(np.prod(shape))
        ptr = self.ctx.alloc_device(nwords * 8)

        class _Raw:
            __cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i8", "data": (ptr, False), "version": 3}

        self._raw = getattr(self, "_raw", []) + [_Raw]
        t = self.torch.as_tensor(_Raw(), device=self.device).view(*shape)
        return t, ptr

    def ipc_export(self, ptr):
        return self.ctx.ipc_export(ptr)

    def ipc_close(self, ptr):
        self.ctx.ipc_close(ptr)

    def free_exportable(self, ptr):
        self._raw = []
        self.ctx.free(ptr)

    def ipc_open(self, handle):
        return self.ctx.ipc_open(handle)

    def lde_scatter(self, coeffs, work_ptr, log_n, log_b, ncols, block_ptrs, block_stride, dup_ptrs, dup_stride):
        from . import GENERATOR
        self.ctx.lde_batch_scatter(coeffs, work_ptr, 1, log_n, log_b, ncols, block_ptrs, block_stride, dup_ptrs, dup_stride,
                                   offset=GENERATOR)

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
