"""prover_mgpu.py — `default_prove` (src/prover.rs:25-174) over G GPUs of one box, one process per GPU.

The reference has no multi-device code (one Metal device, gpu/src/plan.rs:465-469).  What makes the whole prover
shard — not just the first commitment — is the shape of the bit-reversed LDE (src/matrix.rs:225-234): it is
beta = lde_blowup_factor blocks of n rows, block q being the size-n transform of the SAME coefficients over the coset
h_q * <g_n>, h_q = offset * g_N^bitrev(q).  Every step after interpolation is local to a block:

  transform      block q of every column = one size-n coset NTT of the column's coefficients            (LDE)
  leaf hash      a leaf is one LDE row: rows of block q only                                           (src/merkle.rs:412-436)
  constraints    ce-domain point i and its neighbours i + ce_blowup * offset sit in the same block     (src/eval_cpu.rs:115-123)
  DEEP           pointwise over the LDE                                                                (src/composer.rs:89-188)
  FRI fold       a folded value needs ff CONSECUTIVE entries of the bit-reversed codeword               (src/fri.rs:199-231)
  queries        a row and its authentication path below the subtree root live where the row lives

so rank r owns blocks [r * beta/G, (r+1) * beta/G) — a contiguous slab of N/G LDE rows — of every matrix, for the whole
proof.  Only COEFFICIENTS are replicated (n values per column, 1/beta of the LDE): the interpolation of a matrix is
split by columns and the coefficient columns are all-gathered over NVLink (NCCL); no LDE data ever crosses a link.
Per commitment and per FRI layer the G subtree roots (32 bytes each) are all-gathered and every rank finishes the top
log2(G) levels of the tree (src/merkle.rs:485-508), so every rank holds the same transcript: the Fiat–Shamir channel,
the proof-of-work and the query positions are computed identically everywhere and nothing has to be broadcast.  The
proof bytes are identical to the single-GPU prover's (tests/test_gpu_multi.py), hence to the CPU restatement's.

A FRI layer is sharded while every rank still holds at least two of its rows; the remaining small layers are gathered
once and finished on every rank.  torch / torch.distributed are plumbing: buffers, the stream, the collectives.
"""
import hashlib
import time

import numpy as np
import torch

from . import FP, FQ3, GENERATOR as GEN_MONT, ONE
from . import deep
from . import expr as E
from .air import domain_generator
from .channel import ProverChannel
from .proof import FriProof, LayerProof, MerkleView, Proof, Queries
from .prover import GpuProver, ProvingError, _Tree, _canon_rows, _lift, _mont

P = E.P
_R = 2**64
_RINV = pow(_R, -1, P)


def _brev(v, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (v & 1)
        v >>= 1
    return r


def merkle_walk(n_leaves, indices):
    """The index walk of MerkleTreeImpl::prove (src/merkle.rs:149-207; csrc/hash.cu ms_merkle_prove_sha256): which
    leaves and which heap nodes a batched proof names.  Returns (initial leaf indices, sibling leaf indices, node indices)."""
    idx = sorted(set(int(i) for i in indices))
    init, sib, path, node_q = [], [], [], []
    k = 0
    while k < len(idx):
        i = idx[k]
        init.append(i)
        node_q.append((n_leaves + i) >> 1)
        if k + 1 < len(idx) and (i ^ 1) == idx[k + 1]:
            init.append(idx[k + 1])
            k += 2
            continue
        sib.append(i ^ 1)
        k += 1
    head = 0
    while head < len(node_q):
        i = node_q[head]
        head += 1
        if i > 2:
            node_q.append(i >> 1)
        if head < len(node_q) and (i ^ 1) == node_q[head]:
            head += 1
            continue
        path.append(i ^ 1)
    return init, sib, path


def top_levels(sub_roots):
    """heap of the top log2(G) levels from the G subtree roots: top[G + r] = root of rank r's subtree, top[1] = the root"""
    g = len(sub_roots)
    top = [None] * (2 * g)
    for r, d in enumerate(sub_roots):
        top[g + r] = bytes(d)
    for k in range(g - 1, 0, -1):
        top[k] = hashlib.sha256(top[2 * k] + top[2 * k + 1]).digest()
    return top


def node_owner(k, log_g):
    """heap node k of the global tree -> (rank, heap index inside that rank's subtree), or (None, k) for the top levels"""
    d = k.bit_length() - 1
    if d < log_g:
        return None, k
    a = k >> (d - log_g)
    return a - (1 << log_g), (1 << (d - log_g)) + (k - (a << (d - log_g)))


class _ShardedTree:
    """rows [rank * n_local, (rank + 1) * n_local) of a tree with n_total leaves: local leaves / nodes + the shared top"""

    def __init__(self, leaves, nodes, n_local, n_total, top):
        self.leaves, self.nodes, self.n_local, self.n_total, self.top = leaves, nodes, n_local, n_total, top


class _FetchPlan:
    """The query phase needs ~30 LDE rows per matrix, ~30 rows per FRI layer and a few hundred path digests, each living
    on the rank that owns its row.  Requests are collected first (every rank builds the same plan, so every rank knows
    who owns what and how many bytes it is), every rank then fetches what it owns with one device gather per matrix /
    tree, and ONE fixed-layout NCCL all-gather of the packed pieces gives every rank everything."""

    def __init__(self, prover):
        self.p = prover
        self.groups = []          # (fetch(list of local indices) -> sequence of byte strings, [(item, local index)])
        self.items = []           # per handle: (owner rank or None for a literal, nbytes)
        self.literal = {}
        self.data = None

    def _new(self, owner, nbytes):
        self.items.append((owner, nbytes))
        return len(self.items) - 1

    def rows(self, gather_local, n_local, positions, row_words):
        """handles of the rows at global `positions`; gather_local(list of local row ids) -> (k, row_words) array"""
        mine, handles = [], []
        for p in positions:
            owner = p // n_local
            h = self._new(owner, 8 * row_words)
            handles.append(h)
            if owner == self.p.rank:
                mine.append((h, p - owner * n_local))
        if mine:
            self.groups.append((lambda loc, f=gather_local: [np.ascontiguousarray(r, dtype=np.uint64).tobytes() for r in f(loc)], mine))
        return handles

    def view(self, tree, positions):
        """handles of a MerkleView over a sharded tree: (path handles, initial-leaf handles, sibling-leaf handles, height)"""
        init, sib, path = merkle_walk(tree.n_total, positions)
        leaf_mine, node_mine = [], []

        def leaf(i):
            owner, loc = divmod(i, tree.n_local)
            h = self._new(owner, 32)
            if owner == self.p.rank:
                leaf_mine.append((h, loc))
            return h

        def node(k):
            owner, loc = node_owner(k, self.p.log_g)
            h = self._new(owner, 32)
            if owner is None:
                self.literal[h] = tree.top[loc] if loc else bytes(32)     # heap index 0: the unused default digest
            elif owner == self.p.rank:
                node_mine.append((h, loc))
            return h

        hi, hs, hp = [leaf(i) for i in init], [leaf(i) for i in sib], [node(k) for k in path]
        dev = self.p.device

        def fetch(t):
            return lambda loc: [r.tobytes() for r in t.index_select(0, torch.tensor(loc, dtype=torch.int64, device=dev)).cpu().numpy()]

        if leaf_mine:
            self.groups.append((fetch(tree.leaves), leaf_mine))
        if node_mine:
            self.groups.append((fetch(tree.nodes), node_mine))
        return hp, hi, hs, tree.n_total.bit_length() - 1

    def execute(self):
        G, rank = self.p.world, self.p.rank
        mine = {}
        for fn, lst in self.groups:
            for (h, _), v in zip(lst, fn([loc for _, loc in lst])):
                mine[h] = v
        # fixed layout: rank r's buffer = its items in handle order; every rank can compute every offset
        sizes = [0] * G
        offset = {}
        for h, (owner, nbytes) in enumerate(self.items):
            if owner is not None:
                offset[h] = sizes[owner]
                sizes[owner] += nbytes
        cap = max(max(sizes), 1)
        buf = bytearray(cap)
        for h, v in mine.items():
            assert len(v) == self.items[h][1]
            buf[offset[h]:offset[h] + len(v)] = v
        send = torch.frombuffer(buf, dtype=torch.uint8).to(self.p.device)
        recv = torch.empty(G * cap, dtype=torch.uint8, device=self.p.device)
        self.p.dist.all_gather_into_tensor(recv, send)
        raw = recv.cpu().numpy().tobytes()
        self.data = dict(self.literal)
        for h, (owner, nbytes) in enumerate(self.items):
            if owner is not None:
                o = owner * cap + offset[h]
                self.data[h] = raw[o:o + nbytes]
        assert len(self.data) == len(self.items), "a queried row or digest has no owner"

    def get_rows(self, handles):
        return np.frombuffer(b"".join(self.data[h] for h in handles), dtype=np.uint64) if handles else np.zeros(0, dtype=np.uint64)

    def get_view(self, v):
        hp, hi, hs, height = v
        return MerkleView([self.data[h] for h in hp], [self.data[h] for h in hi], [self.data[h] for h in hs], height)


class ShardedProver(GpuProver):
    def __init__(self, dist, device):
        super().__init__(device)
        self.dist = dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world & (self.world - 1):
            raise ValueError("world size must be a power of two")
        self.log_g = self.world.bit_length() - 1

    # ---- collectives (on the prover's stream: torch orders NCCL against the current stream)
    def _all_gather_bytes(self, b):
        t = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(self.device)
        out = torch.empty(self.world * len(b), dtype=torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(out, t)
        raw = out.cpu().numpy().tobytes()
        return [raw[i * len(b):(i + 1) * len(b)] for i in range(self.world)]

    def _all_gather_objects(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    # ---- building blocks
    def _interpolate(self, evals, field, ncols, log_n, from_host=False):
        """Matrix::interpolate (src/matrix.rs:101-116): the inverse transforms are split by columns, the coefficient
        columns all-gathered — every rank ends up with the whole (ncols, n) coefficient matrix.  from_host: `evals` is a
        host matrix of which only this rank's column block is uploaded."""
        n, G = 1 << log_n, self.world
        if ncols < G:
            polys = self._empty(ncols, n * field)
            self.ctx.ntt_batch_to(evals, polys, field, log_n, ncols, inverse=True)
            return polys
        per = (ncols + G - 1) // G
        pad = self._empty(G * per, n * field)
        lo, hi = min(self.rank * per, ncols), min((self.rank + 1) * per, ncols)
        if hi > lo:
            if from_host:
                pad[lo:hi].copy_(self._to_device(evals[lo:hi]))
                self.ctx.ntt_batch(pad[lo], field, log_n, hi - lo, inverse=True)
            else:
                self.ctx.ntt_batch_to(evals[lo], pad[lo], field, log_n, hi - lo, inverse=True)
        self.dist.all_gather_into_tensor(pad.view(-1), pad[self.rank * per:(self.rank + 1) * per].reshape(-1))
        return pad[:ncols]

    def _offsets(self, log_n, log_b):
        """Montgomery words of h_q = offset * g_N^bitrev(q) for this rank's blocks"""
        gN = domain_generator(log_n + log_b)
        bpr = (1 << log_b) // self.world
        return [(q, 7 * pow(gN, _brev(q, log_b), P) % P * _R % P) for q in range(self.rank * bpr, (self.rank + 1) * bpr)]

    def _lde_slab(self, polys, field, ncols, log_n, log_b):
        """this rank's blocks of the bit-reversed LDE of every column: (ncols, N/G) elements, block j at rows [j n, (j+1) n)"""
        n = 1 << log_n
        rows_per = (n << log_b) // self.world
        slab = self._empty(ncols, rows_per * field)
        for j, (_, h) in enumerate(self._offsets(log_n, log_b)):
            self.ctx.lde_batch(polys, slab.data_ptr() + j * n * field * 8, field, log_n, 0, ncols, in_stride=n,
                               out_stride=rows_per, offset=h, bitrev=True)
        return slab

    def _finish_tree(self, sub_root, leaves, nodes, n_local):
        top = top_levels(self._all_gather_bytes(sub_root))
        return _ShardedTree(leaves, nodes, n_local, n_local * self.world, top), top[1]

    def _commit_slab(self, slab, field, ncols, rows_per):
        leaves, nodes = self._empty(rows_per, 4), self._empty(rows_per, 4)
        sub = self.ctx.merkle_commit(slab, field, rows_per, ncols, col_stride=rows_per, leaves=leaves, nodes=nodes)
        return self._finish_tree(sub, leaves, nodes, rows_per)

    def _block_ptrs(self, slab, field, ncols, rows_per, j, n):
        return [slab.data_ptr() + (c * rows_per + j * n) * field * 8 for c in range(ncols)]

    def fri_commit(self, slab, log_n, fq, options, channel):
        """FriProver::build_layers (src/fri.rs:199-231) on a codeword sharded by rows: `slab` holds this rank's
        2^log_n / G consecutive entries of the bit-reversed codeword.  Per layer: rows of ff entries -> leaf hashes ->
        subtree -> all-gather of the G subtree roots -> channel; the fold is local (row k needs only row k).  Layers with
        fewer than 2 rows per rank are gathered once and finished on every rank.  Returns (layers, remainder codeword
        (replicated), its log size); layers = [(evals, tree, root, rows in the layer, sharded?)]."""
        ctx, dist, G, rank = self.ctx, self.dist, self.world, self.rank
        ff = options.fri_folding_factor
        log_ff = ff.bit_length() - 1
        layers = []
        cur, ln, sharded = slab, log_n, True
        for _ in range(options.fri_num_layers(1 << log_n)):
            nrows = 1 << (ln - log_ff)
            if sharded and nrows // G < 2:
                full = self._empty((1 << ln) * fq)
                dist.all_gather_into_tensor(full, cur)
                cur, sharded = full, False
            if sharded:
                nloc = nrows // G
                leaves, nodes = self._empty(nloc, 4), self._empty(nloc, 4)
                sub = ctx.merkle_commit_rows(cur, ff * fq, nloc, leaves=leaves, nodes=nodes)
                tree, root = self._finish_tree(sub, leaves, nodes, nloc)
                channel.commit_fri_layer(root)
                layers.append((cur, tree, root, nrows, True))
                alpha = channel.draw_fri_alpha()
                nxt = self._empty(nloc * fq)
                off = pow(domain_generator(ln), _brev(rank, self.log_g), P) * _R % P      # ONE * g_(2^ln)^bitrev(rank)
                ctx.fri_fold(cur, nxt, fq, ln - self.log_g, log_ff,
                             np.array([_mont(c) for c in _lift(alpha)], dtype=np.uint64), offset=off)
            else:
                leaves, nodes = self._empty(nrows, 4), self._empty(nrows, 4)
                root = ctx.merkle_commit_rows(cur, ff * fq, nrows, leaves=leaves, nodes=nodes)
                channel.commit_fri_layer(root)
                layers.append((cur, _Tree(leaves, nodes, nrows), root, nrows, False))
                alpha = channel.draw_fri_alpha()
                nxt = self._empty(nrows * fq)
                ctx.fri_fold(cur, nxt, fq, ln, log_ff, np.array([_mont(c) for c in _lift(alpha)], dtype=np.uint64))
            cur, ln = nxt, ln - log_ff
        if sharded:
            full = self._empty((1 << ln) * fq)
            dist.all_gather_into_tensor(full, cur)
            cur = full
        return layers, cur, ln

    # ---- default_prove
    def _prove(self, stark, options, witness):
        ctx, dist, G, rank = self.ctx, self.dist, self.world, self.rank
        cfg = stark.AirConfig
        timings = {}
        t_all = t0 = time.perf_counter()

        # NVTX range per prover phase (visible to nsys / ncu --nvtx): the range of phase k is closed and the range of
        # phase k + 1 opened where the reference prints its per-phase timings (src/prover.rs:40-170)
        phases = ["init_air", "base_trace_commitment", "extension_trace_commitment", "constraint_eval",
                  "composition_trace_commitment", "deep_composition", "fri", "proof_of_work", "queries"]
        torch.cuda.nvtx.range_push("prove:" + phases[0])

        def lap(name):
            nonlocal t0
            ctx.sync()
            t = time.perf_counter()
            timings[name] = t - t0
            t0 = t
            torch.cuda.nvtx.range_pop()
            k = phases.index(name) + 1
            if k < len(phases):
                torch.cuda.nvtx.range_push("prove:" + phases[k])

        import copy
        from .air import Air
        trace = stark.generate_trace(witness)
        n = len(trace)
        key = (cfg, n, options)
        if key not in self._airs:
            air0 = Air(cfg, n, None, options)
            air0.composition_program()
            air0.deep_program()
            # the composition evaluated block by block: inside a block the ce-domain stride is 1 and the domain has n points
            air0._block_program = E.compile_program(air0.composition_constraint, cfg.NUM_BASE_COLUMNS, lde_step=1,
                                                    log_ce=air0.log_n, symbolic=True, batch_inverses=True)
            air0.num_challenges(), air0.num_composition_constraint_coeffs(), air0.trace_arguments()
            self._airs[key] = air0
        air = copy.copy(self._airs[key])
        air.public_inputs = stark.get_public_inputs()
        channel = ProverChannel(air, stark.gen_public_coin(air), ctx)
        fq = FP if cfg.FQ_IS_FP else FQ3
        log_n = air.log_n
        beta = options.lde_blowup_factor
        log_b = beta.bit_length() - 1
        log_N, N = log_n + log_b, n * beta
        if beta % G or n < 16:
            raise ProvingError("the sharded prover needs a world size dividing the LDE blow-up factor and n >= 16")
        bpr, rows_per = beta // G, N // G
        my_blocks = self._offsets(log_n, log_b)
        nbase, next_ = cfg.NUM_BASE_COLUMNS, cfg.NUM_EXTENSION_COLUMNS
        lap("init_air")

        # ---- base trace commitment (prover.rs:46-55)
        host_base = trace.base_columns()
        if tuple(host_base.shape) != (nbase, n):
            raise ProvingError(f"expected {nbase} base columns of {n} rows")
        needs_full_base = next_ > 0          # extension columns are built from the whole base trace (on every rank)
        if needs_full_base or nbase < G or (isinstance(host_base, torch.Tensor) and host_base.is_cuda):
            base = self._to_device(host_base)
            base_polys = self._interpolate(base, FP, nbase, log_n)
        else:
            base = None
            base_polys = self._interpolate(host_base, FP, nbase, log_n, from_host=True)
        base_slab = self._lde_slab(base_polys, FP, nbase, log_n, log_b)
        base_tree, base_root = self._commit_slab(base_slab, FP, nbase, rows_per)
        channel.commit_base_trace(base_root)
        lap("base_trace_commitment")
        challenges = [channel.public_coin.draw() for _ in range(air.num_challenges())]
        hints = air.gen_hints(challenges)

        # ---- extension trace commitment (prover.rs:56-72): built from the (replicated) base trace on every rank
        if hasattr(trace, "build_extension_columns_device"):
            ext = trace.build_extension_columns_device(challenges, ctx, base)
        else:
            ext = trace.build_extension_columns(challenges)
        del base
        num_ext = 0 if ext is None else int(ext.shape[0])
        if num_ext != next_:
            raise ProvingError(f"expected {next_} extension columns, got {num_ext}")
        ext_polys = ext_slab = ext_tree = None
        if ext is not None:
            ext_polys = self._interpolate(self._to_device(ext), fq, next_, log_n)
            ext_slab = self._lde_slab(ext_polys, fq, next_, log_n, log_b)
            ext_tree, ext_root = self._commit_slab(ext_slab, fq, next_, rows_per)
            channel.commit_extension_trace(ext_root)
        del ext
        lap("extension_trace_commitment")

        # ---- constraint evaluation (prover.rs:75-108), block by block: the blocks q < ce_blowup of the LDE are the ce
        # domain; each is evaluated where it lives, then the (small) evaluation column is shared
        ce_blowup = air.ce_blowup_factor
        log_ce = log_n + ce_blowup.bit_length() - 1
        M = n * ce_blowup
        composition_coeffs = [channel.public_coin.draw() for _ in range(air.num_composition_constraint_coeffs())]
        prog = air._block_program.bind(challenges=challenges, hints=hints, ccoefs=composition_coeffs)
        comp_evals = self._empty(M * fq)
        is_fq = [False] * nbase + [True] * next_
        for j, (q, h) in enumerate(my_blocks):
            if q < ce_blowup:
                cols = self._block_ptrs(base_slab, FP, nbase, rows_per, j, n)
                if next_:
                    cols += self._block_ptrs(ext_slab, fq, next_, rows_per, j, n)
                ctx.eval_constraints_ptrs(prog, comp_evals[q * n * fq:(q + 1) * n * fq], log_n, cols, is_fq, fq_field=fq,
                                          offset=h, trace_bitrev=True, out_bitrev=True)
        for q in range(ce_blowup):
            dist.broadcast(comp_evals[q * n * fq:(q + 1) * n * fq], src=q // bpr)
        lap("constraint_eval")

        # ---- composition trace (prover.rs:110-125): one column over the ce coset -> coefficients -> ce_blowup columns.
        # comp_evals is the bit-reversed ce-domain column; one small transform, done on every rank
        ctx.bit_reverse(comp_evals, fq, log_ce)
        ctx.ntt_batch(comp_evals, fq, log_ce, 1, inverse=True, offset=GEN_MONT)
        if ce_blowup == 1:
            comp_polys = comp_evals.view(1, n * fq)
        else:
            comp_polys = self._empty(ce_blowup, n * fq)
            ctx.matrix_from_rows(comp_evals, comp_polys, fq, n, ce_blowup)
        comp_slab = self._lde_slab(comp_polys, fq, ce_blowup, log_n, log_b)
        comp_tree, comp_root = self._commit_slab(comp_slab, fq, ce_blowup, rows_per)
        channel.commit_composition_trace(comp_root)
        lap("composition_trace_commitment")

        # ---- out-of-domain evaluations (composer.rs:43-86) from the replicated coefficients
        z = channel.get_ood_point()
        zq = _lift(z)
        trace_arguments = air.trace_arguments()
        offsets = sorted(set(o for _, o in trace_arguments))
        z_points, z_m = deep.ood_points(zq, log_n, offsets, ce_blowup)
        pts = np.array([[_mont(c) for c in z_points[o]] for o in offsets], dtype=np.uint64).reshape(-1, 3)
        base_ood = ctx.poly_eval(base_polys, FP, n, nbase, pts)
        ext_ood = ctx.poly_eval(ext_polys, fq, n, next_, pts) if next_ else None
        comp_ood = ctx.poly_eval(comp_polys, fq, n, ce_blowup, np.array([[_mont(c) for c in z_m]], dtype=np.uint64))

        def unlift(w3):
            t = tuple(int(w) * _RINV % P for w in w3)
            if fq == FP:
                if t[1] or t[2]:
                    raise ProvingError("out-of-domain value left the base field although Fq = Fp")
                return t[0]
            return t

        execution_trace_oods = []
        for col, off in trace_arguments:
            k = offsets.index(off)
            if col < nbase:
                execution_trace_oods.append(unlift(base_ood[col, k]))
            elif col < nbase + next_:
                execution_trace_oods.append(unlift(ext_ood[col - nbase, k]))
            else:
                raise ProvingError(f"column is {col} but there are only {nbase + next_} columns")
        composition_trace_oods = [unlift(comp_ood[j, 0]) for j in range(ce_blowup)]
        channel.send_ood_evals(execution_trace_oods, composition_trace_oods)

        # ---- DEEP composition polynomial over this rank's LDE rows (composer.rs:89-188 in evaluation form)
        ex_alphas, co_alphas, (d_alpha, d_beta) = stark.gen_deep_coeffs(channel.public_coin, air)
        dprog_sym, dkeys = air.deep_program()
        dprog = dprog_sym.bind(hints=deep.deep_hint_values(
            dkeys, z_points, z_m, [_lift(v) for v in execution_trace_oods], [_lift(v) for v in composition_trace_oods],
            [_lift(v) for v in ex_alphas], [_lift(v) for v in co_alphas], _lift(d_alpha), _lift(d_beta),
            trace_arguments=trace_arguments))
        ncols_all = nbase + next_ + ce_blowup
        deep_slab = self._empty(rows_per * fq)
        for j, (q, h) in enumerate(my_blocks):
            cols = self._block_ptrs(base_slab, FP, nbase, rows_per, j, n)
            if next_:
                cols += self._block_ptrs(ext_slab, fq, next_, rows_per, j, n)
            cols += self._block_ptrs(comp_slab, fq, ce_blowup, rows_per, j, n)
            ctx.eval_constraints_ptrs(dprog, deep_slab[j * n * fq:(j + 1) * n * fq], log_n, cols,
                                      [False] * nbase + [True] * (ncols_all - nbase), fq_field=fq, offset=h,
                                      trace_bitrev=True, out_bitrev=True)
        lap("deep_composition")

        # ---- FRI (fri.rs:179-249): layers sharded by rows while every rank keeps >= 2 rows of the layer
        layers, cur, ln = self.fri_commit(deep_slab, log_N, fq, options, channel)
        ff = options.fri_folding_factor
        rem_size = 1 << ln
        if rem_size > options.fri_max_remainder_coeffs * beta:
            raise ProvingError("remainder domain too large")
        rem = cur.clone()
        ctx.bit_reverse(rem, fq, ln)
        ctx.ntt_batch(rem, fq, ln, 1, inverse=True, offset=ONE)
        ctx.sync()
        rem_coeffs = _canon_rows(rem.cpu().numpy().view(np.uint64), fq)
        keep = rem_size // beta
        zero = 0 if fq == FP else (0, 0, 0)
        if any(c != zero for c in rem_coeffs[keep:]):
            raise ProvingError("FRI remainder is not low degree: the trace does not satisfy the AIR (fri.rs:246)")
        channel.commit_remainder(rem_coeffs[:keep])
        lap("fri")

        channel.grind_fri_commitments()
        lap("proof_of_work")

        # ---- queries (fri.rs:151-177, trace.rs:115-157): rows and path digests come from the ranks that own them
        positions = channel.get_fri_query_positions()
        pos_sorted = sorted(set(positions))
        plan = _FetchPlan(self)
        pending, folded = [], positions
        for evals, tree, root, nrows, was_sharded in layers:
            folded = sorted(set(p // ff for p in folded))
            if was_sharded:
                nloc = nrows // G
                hr = plan.rows(lambda loc, e=evals, k=nloc: ctx.gather_rows_rowmajor(e, ff * fq, k, loc), nloc, folded, ff * fq)
                pending.append((root, hr, plan.view(tree, folded), None))
            else:
                rows = ctx.gather_rows_rowmajor(evals, ff * fq, nrows, folded)
                pending.append((root, None, None, (rows, self._view(tree, folded))))

        def trace_rows(slab, field, ncols):
            # Queries::new keeps the caller's position order (sorted, deduplicated by draw_queries)
            return plan.rows(lambda loc: ctx.gather_rows(slab, field, rows_per, ncols, loc, col_stride=rows_per), rows_per, positions,
                             ncols * field)

        h_base, h_comp = trace_rows(base_slab, FP, nbase), trace_rows(comp_slab, fq, ce_blowup)
        h_ext = trace_rows(ext_slab, fq, next_) if next_ else None
        v_base, v_comp = plan.view(base_tree, pos_sorted), plan.view(comp_tree, pos_sorted)
        v_ext = plan.view(ext_tree, pos_sorted) if next_ else None
        plan.execute()
        fri_layers = []
        for root, hr, hv, direct in pending:
            rows, view = (plan.get_rows(hr), plan.get_view(hv)) if direct is None else direct
            fri_layers.append(LayerProof(_canon_rows(rows, fq), view, root))
        fri_proof = FriProof(fri_layers, channel.fri_remainder_coeffs)
        queries = Queries(
            _canon_rows(plan.get_rows(h_base), 1),
            _canon_rows(plan.get_rows(h_ext), fq) if next_ else [],
            _canon_rows(plan.get_rows(h_comp), fq),
            plan.get_view(v_base), plan.get_view(v_ext) if next_ else None, plan.get_view(v_comp))
        lap("queries")
        timings["total"] = time.perf_counter() - t_all
        return Proof(options, n, channel.base_trace_commitment, channel.extension_trace_commitment,
                     channel.composition_trace_commitment, fri_proof, channel.pow_nonce, queries,
                     channel.execution_trace_ood_evals, channel.composition_trace_ood_evals, timings)
