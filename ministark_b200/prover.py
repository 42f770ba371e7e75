"""prover.py — `default_prove` (src/prover.rs:25-174) with every data-parallel step on the B200.

Same order of commitments and Fiat–Shamir draws as the reference, so the transcript — and with it every
root, out-of-domain value, FRI layer and query in the proof — is determined by the same inputs.  What changes is
where the data lives and how each step is computed:

  reference (per step, host <-> unified memory)            here (resident in HBM, only digests/rows return)
  ---------------------------------------------------------------------------------------------------------
  Matrix::interpolate, bit_reversed_evaluate               ms_ntt_batch_to + ms_lde_batch (bit-reversed out)
  MerkleTree::from_matrix (CPU SHA-256)                    ms_merkle_commit_sha256, leaves + nodes stay on device
  bit_reverse_ce_trace x2 + eval_cpu::eval                 ms_eval_constraints on the bit-reversed LDE prefix
  into_polynomials + push loop + LDE                       ms_ntt_batch (ce coset) + ms_matrix_from_rows + ms_lde_batch
  horner_evaluate per column                               ms_poly_eval
  divide_out_points_into, sum_columns, degree adjust, LDE  one fused pointwise evaluation over the LDE (deep.py)
  apply_drp (iNTT + fold + NTT + 2 bit reversals)          ms_merkle_commit_rows_sha256 + ms_fri_fold per layer
  grind_proof_of_work (rayon find_any)                     ms_pow_grind_sha256 (smallest nonce)
  prove_rows / get_row                                     ms_merkle_prove_sha256 + ms_gather_rows(_rowmajor)

torch is used for device buffers and the stream only.  There is no CPU fallback: without the CUDA library the
Context constructor raises.
"""
import copy
import time

import numpy as np
import torch

from . import FP, FQ3, GENERATOR as GEN_MONT, ONE, Context
from . import deep
from . import expr as E
from .air import Air
from .channel import ProverChannel, PublicCoin, serialize_element
from .proof import FriProof, LayerProof, MerkleView, Proof, Queries

P = E.P
_R = 2**64
_RINV = pow(_R, -1, P)


class ProvingError(RuntimeError):
    pass


def _mont(v):
    return int(v) * _R % P


def _canon_rows(words, fq_words):
    """numpy Montgomery words -> flat list of canonical ints (fq_words == 1) or 3-tuples"""
    flat = [int(w) * _RINV % P for w in np.asarray(words, dtype=np.uint64).ravel()]
    if fq_words == 1:
        return flat
    return [tuple(flat[i:i + 3]) for i in range(0, len(flat), 3)]


def _lift(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (int(v), 0, 0)


class Trace:
    """src/trace.rs:13-35.  base_columns(): (ncols, n) uint64 Montgomery words — a numpy array or a cuda int64 tensor.
    build_extension_columns(challenges): None, or (ncols, n * fq_words) words of Fq elements."""

    def __init__(self, base, extension_builder=None):
        self._base, self._ext = base, extension_builder

    def base_columns(self):
        return self._base

    def __len__(self):
        return int(self._base.shape[1])

    def build_extension_columns(self, challenges):
        return None if self._ext is None else self._ext(challenges)


class _Tree:
    def __init__(self, leaves, nodes, n):
        self.leaves, self.nodes, self.n = leaves, nodes, n


class Stark:
    """src/stark.rs:24-85.  Subclass: set AirConfig, implement get_public_inputs / generate_trace and, if the
    public inputs are not a single Fp, public_inputs_bytes()."""
    AirConfig = None

    def get_public_inputs(self):
        raise NotImplementedError

    def generate_trace(self, witness):
        return witness

    def public_inputs_bytes(self, public_inputs):
        """CanonicalSerialize of PublicInputs (compressed).  Default: one field element, or a tuple/list of them
        (ark-serialize writes tuple members back to back)."""
        if isinstance(public_inputs, list):
            return b"".join(serialize_element(v) for v in public_inputs)
        return serialize_element(public_inputs)

    def gen_public_coin(self, air):
        """examples/fib/main.rs:166-172: SHA-256(public inputs ‖ trace_len ‖ options), all serialize_compressed"""
        import hashlib
        seed = self.public_inputs_bytes(air.public_inputs) + int(air.trace_len).to_bytes(8, "little") + air.options.to_bytes()
        return PublicCoin(hashlib.sha256(seed).digest(), ext=not self.AirConfig.FQ_IS_FP)

    def gen_deep_coeffs(self, public_coin, air):
        """src/stark.rs:42-54"""
        ex = [public_coin.draw() for _ in range(len(air.trace_arguments()))]
        co = [public_coin.draw() for _ in range(air.ce_blowup_factor)]
        return ex, co, (public_coin.draw(), public_coin.draw())

    def prove(self, options, witness, device=0):
        """Stark::prove (src/stark.rs:57-63).  The per-device prover (context, stream, compiled AIR programs) is created
        on first use and reused, like the reference's process-global Planner."""
        return GpuProver.shared(device).prove(self, options, witness)


class GpuProver:
    """owns the device context (one in-order stream) and runs default_prove"""
    _shared = {}

    @classmethod
    def shared(cls, device=0):
        if device not in cls._shared:
            cls._shared[device] = cls(device)
        return cls._shared[device]

    def __init__(self, device=0):
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.ctx = Context(device, stream=self.stream.cuda_stream)
        self._airs = {}

    # ---- helpers
    def _to_device(self, a):
        if isinstance(a, torch.Tensor):
            return a.to(self.device)
        a = np.ascontiguousarray(a, dtype=np.uint64)
        return torch.from_numpy(a.view(np.int64)).to(self.device, non_blocking=False)

    def _empty(self, *shape):
        return torch.empty(shape, dtype=torch.int64, device=self.device)

    def _commit_columns(self, polys_in, field, log_n, log_b, ncols, is_evals):
        """interpolate (if is_evals) + bit-reversed LDE + Merkle commit of a column-major matrix.
        Returns (polys, lde, tree, root)."""
        ctx, n, N = self.ctx, 1 << log_n, 1 << (log_n + log_b)
        if is_evals:
            polys = self._empty(ncols, n * field)
            ctx.ntt_batch_to(polys_in, polys, field, log_n, ncols, inverse=True)      # Matrix::interpolate over the trace domain
        else:
            polys = polys_in
        lde = self._empty(ncols, N * field)
        ctx.lde_batch(polys, lde, field, log_n, log_b, ncols, offset=GEN_MONT, bitrev=True)
        leaves, nodes = self._empty(N, 4), self._empty(N, 4)
        root = ctx.merkle_commit(lde, field, N, ncols, leaves=leaves, nodes=nodes)
        return polys, lde, _Tree(leaves, nodes, N), root

    def _view(self, tree, positions):
        nodes, init, sib, height = self.ctx.merkle_prove(tree.leaves, tree.nodes, tree.n, positions)
        return MerkleView(nodes, init, sib, height)

    # ---- default_prove
    def prove(self, stark, options, witness):
        with torch.cuda.stream(self.stream):
            return self._prove(stark, options, witness)

    def _prove(self, stark, options, witness):
        ctx = self.ctx
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

        trace = stark.generate_trace(witness)
        n = len(trace)
        # the AIR bookkeeping and its two compiled evaluator programs depend on (AirConfig, trace length, options) only:
        # built once per prover, then shared by every proof (public inputs and verifier randomness are bound per proof)
        key = (cfg, n, options)
        if key not in self._airs:
            self._airs[key] = Air(cfg, n, None, options)
            self._airs[key].composition_program()
            self._airs[key].deep_program()
            self._airs[key].num_challenges(), self._airs[key].num_composition_constraint_coeffs(), self._airs[key].trace_arguments()
        air = copy.copy(self._airs[key])
        air.public_inputs = stark.get_public_inputs()
        channel = ProverChannel(air, stark.gen_public_coin(air), ctx)
        fq = FP if cfg.FQ_IS_FP else FQ3
        log_n = air.log_n
        beta = options.lde_blowup_factor
        log_b = beta.bit_length() - 1
        log_N, N = log_n + log_b, n * beta
        nbase, next_ = cfg.NUM_BASE_COLUMNS, cfg.NUM_EXTENSION_COLUMNS
        lap("init_air")

        # ---- base trace commitment (prover.rs:46-55).  A host trace is uploaded in column chunks on a second stream
        # while the previous chunk is interpolated and extended (columns are independent until the row hash); a
        # pinned trace — the analogue of the reference's GpuAllocator-backed columns — makes the copies asynchronous.
        host_base = trace.base_columns()
        if tuple(host_base.shape) != (nbase, n):
            raise ProvingError(f"expected {nbase} base columns of {n} rows")
        if isinstance(host_base, torch.Tensor) and host_base.is_cuda:
            base = host_base.to(self.device)
            base_polys, base_lde, base_tree, base_root = self._commit_columns(base, FP, log_n, log_b, nbase, True)
        else:
            if not isinstance(host_base, torch.Tensor):
                host_base = torch.from_numpy(np.ascontiguousarray(host_base, dtype=np.uint64).view(np.int64))
            base, base_polys, base_lde = self._empty(nbase, n), self._empty(nbase, n), self._empty(nbase, N)
            chunk = max(1, min(nbase, (64 << 20) // (8 * n) or 1))            # ~64 MiB per copy
            self.copy_stream.wait_stream(self.stream)
            events = []
            with torch.cuda.stream(self.copy_stream):
                for c0 in range(0, nbase, chunk):
                    c1 = min(c0 + chunk, nbase)
                    base[c0:c1].copy_(host_base[c0:c1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.copy_stream)
                    events.append((c0, c1, ev))
            for c0, c1, ev in events:
                self.stream.wait_event(ev)
                ctx.ntt_batch_to(base[c0], base_polys[c0], FP, log_n, c1 - c0, inverse=True)
                ctx.lde_batch(base_polys[c0], base_lde[c0], FP, log_n, log_b, c1 - c0, offset=GEN_MONT, bitrev=True)
            leaves, nodes = self._empty(N, 4), self._empty(N, 4)
            base_root = ctx.merkle_commit(base_lde, FP, N, nbase, leaves=leaves, nodes=nodes)
            base_tree = _Tree(leaves, nodes, N)
        channel.commit_base_trace(base_root)
        lap("base_trace_commitment")
        challenges = [channel.public_coin.draw() for _ in range(air.num_challenges())]
        hints = air.gen_hints(challenges)

        # ---- extension trace commitment (prover.rs:56-72)
        if hasattr(trace, "build_extension_columns_device"):
            # running products / evaluations as device scans over the resident base trace
            ext = trace.build_extension_columns_device(challenges, ctx, base)
        else:
            ext = trace.build_extension_columns(challenges)
        del base
        num_ext = 0 if ext is None else int(ext.shape[0])
        if num_ext != next_:
            raise ProvingError(f"expected {next_} extension columns, got {num_ext}")
        ext_polys = ext_lde = ext_tree = None
        if ext is not None:
            ext_polys, ext_lde, ext_tree, ext_root = self._commit_columns(self._to_device(ext), fq, log_n, log_b, next_, True)
            channel.commit_extension_trace(ext_root)
        del ext
        lap("extension_trace_commitment")

        # ---- constraint evaluation over the ce domain (prover.rs:75-108).  The first M entries of a bit-reversed LDE
        # column ARE the ce-coset evaluations in bit-reversed order, so they are read in place (trace_bitrev).
        ce_blowup = air.ce_blowup_factor
        log_ce = log_n + ce_blowup.bit_length() - 1
        M = n * ce_blowup
        composition_coeffs = [channel.public_coin.draw() for _ in range(air.num_composition_constraint_coeffs())]
        prog = air.composition_program().bind(challenges=challenges, hints=hints, ccoefs=composition_coeffs)
        comp_evals = self._empty(M * fq)
        ctx.eval_constraints(prog, comp_evals, log_ce, base_cols=base_lde, nbase=nbase, base_stride=N,
                             ext_cols=ext_lde, next_=next_, ext_stride=N, fq_field=fq, offset=GEN_MONT, trace_bitrev=True)
        lap("constraint_eval")

        # ---- composition trace (prover.rs:110-125): coefficients over the ce coset, column i = coefficients = i mod ce_blowup
        ctx.ntt_batch(comp_evals, fq, log_ce, 1, inverse=True, offset=GEN_MONT)
        if ce_blowup == 1:
            comp_polys = comp_evals.view(1, n * fq)
        else:
            comp_polys = self._empty(ce_blowup, n * fq)
            ctx.matrix_from_rows(comp_evals, comp_polys, fq, n, ce_blowup)
        _, comp_lde, comp_tree, comp_root = self._commit_columns(comp_polys, fq, log_n, log_b, ce_blowup, False)
        channel.commit_composition_trace(comp_root)
        lap("composition_trace_commitment")

        # ---- out-of-domain evaluations (composer.rs:43-86)
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

        # ---- DEEP composition polynomial, evaluated straight over the LDE domain (composer.rs:89-188 in evaluation form)
        ex_alphas, co_alphas, (d_alpha, d_beta) = stark.gen_deep_coeffs(channel.public_coin, air)
        dprog_sym, dkeys = air.deep_program()
        dprog = dprog_sym.bind(hints=deep.deep_hint_values(
            dkeys, z_points, z_m, [_lift(v) for v in execution_trace_oods], [_lift(v) for v in composition_trace_oods],
            [_lift(v) for v in ex_alphas], [_lift(v) for v in co_alphas], _lift(d_alpha), _lift(d_beta),
            trace_arguments=trace_arguments))
        ncols_all = nbase + next_ + ce_blowup
        sz = N * 8
        cols = [base_lde.data_ptr() + c * sz for c in range(nbase)]
        cols += [ext_lde.data_ptr() + c * sz * fq for c in range(next_)]
        cols += [comp_lde.data_ptr() + c * sz * fq for c in range(ce_blowup)]
        deep_lde = self._empty(N * fq)
        ctx.eval_constraints_ptrs(dprog, deep_lde, log_N, cols, [False] * nbase + [True] * (ncols_all - nbase), fq_field=fq,
                                  offset=GEN_MONT, trace_bitrev=True, out_bitrev=True)
        lap("deep_composition")

        # ---- FRI (fri.rs:179-249)
        ff = options.fri_folding_factor
        log_ff = ff.bit_length() - 1
        layers = []
        cur, ln = deep_lde, log_N
        for _ in range(options.fri_num_layers(N)):
            nrows = 1 << (ln - log_ff)
            leaves, nodes = self._empty(nrows, 4), self._empty(nrows, 4)
            root = ctx.merkle_commit_rows(cur, ff * fq, nrows, leaves=leaves, nodes=nodes)   # Matrix::from_arrays + from_matrix
            channel.commit_fri_layer(root)
            layers.append((cur, _Tree(leaves, nodes, nrows), root, nrows))
            alpha = channel.draw_fri_alpha()
            nxt = self._empty(nrows * fq)
            ctx.fri_fold(cur, nxt, fq, ln, log_ff, np.array([_mont(c) for c in _lift(alpha)], dtype=np.uint64))   # apply_drp, offset ONE
            cur, ln = nxt, ln - log_ff
        # set_remainder (fri.rs:233-249)
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

        # ---- queries (fri.rs:151-177, trace.rs:115-157)
        positions = channel.get_fri_query_positions()
        fri_layers, folded = [], positions
        for evals, tree, root, nrows in layers:
            folded = sorted(set(p // ff for p in folded))                                    # fold_positions
            rows = ctx.gather_rows_rowmajor(evals, ff * fq, nrows, folded)
            fri_layers.append(LayerProof(_canon_rows(rows, fq), self._view(tree, folded), root))
        fri_proof = FriProof(fri_layers, channel.fri_remainder_coeffs)
        queries = Queries(
            _canon_rows(ctx.gather_rows(base_lde, FP, N, nbase, positions), 1),
            _canon_rows(ctx.gather_rows(ext_lde, fq, N, next_, positions), fq) if next_ else [],
            _canon_rows(ctx.gather_rows(comp_lde, fq, N, ce_blowup, positions), fq),
            self._view(base_tree, positions),
            self._view(ext_tree, positions) if next_ else None,
            self._view(comp_tree, positions))
        lap("queries")
        timings["total"] = time.perf_counter() - t_all
        return Proof(options, n, channel.base_trace_commitment, channel.extension_trace_commitment,
                     channel.composition_trace_commitment, fri_proof, channel.pow_nonce, queries,
                     channel.execution_trace_ood_evals, channel.composition_trace_ood_evals, timings)
