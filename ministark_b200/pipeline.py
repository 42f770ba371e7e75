"""pipeline.py — the base-trace commitment + constraint-evaluation leg of `default_prove`
(src/prover.rs:46-55 and :86-108) as one resident pipeline on a B200:

    host trace (pinned) --H2D--> iNTT --> coset LDE (bit-reversed) --> SHA-256 Merkle commit
                                                   \\--> fused constraint evaluation on the LDE prefix

The reference does this with `Matrix::interpolate`, `bit_reversed_evaluate`, `MerkleTree::from_matrix`,
two `bit_reverse_ce_trace` calls and `eval_constraint`, bouncing through host memory between every
step (unified memory on Apple).  Here only the trace goes up and only the root and the composition
evaluations come back.  The upload is chunked by columns and overlapped with the transforms of the
previous chunk (CUDA streams + events): columns are independent until the row hash.

torch is used for device buffers, streams and events only.
"""
import torch

from . import FP, GENERATOR


class TraceCommitPipeline:
    def __init__(self, ctx, device, log_n, log_blowup, ncols, evaluator=None, chunk_cols=4, stream=None, lde_fn=None):
        self.ctx, self.device = ctx, device
        self.log_n, self.log_b, self.ncols = log_n, log_blowup, ncols
        self.n, self.N = 1 << log_n, 1 << (log_n + log_blowup)
        self.chunk = max(1, min(chunk_cols, ncols))
        self.evaluator = evaluator
        self.lde_fn = lde_fn          # multi-GPU: ShardedCommit.lde_columns (LDE fused with the exchange into row slabs)
        self.compute = stream if stream is not None else torch.cuda.current_stream(device)
        self.copy = torch.cuda.Stream(device=device)
        i64 = torch.int64
        self.trace = torch.empty((ncols, self.n), dtype=i64, device=device)
        self.polys = torch.empty((ncols, self.n), dtype=i64, device=device)
        self.lde = torch.empty((ncols, self.N), dtype=i64, device=device)
        self.leaves = torch.empty((self.N, 4), dtype=i64, device=device)
        self.nodes = torch.empty((self.N, 4), dtype=i64, device=device)
        self.ce = torch.empty(self.n, dtype=i64, device=device) if evaluator is not None else None
        self.host_ce = torch.empty(self.n, dtype=i64, pin_memory=True) if evaluator is not None else None

    # ---- resident stages (inputs already in HBM)
    def transform(self, c0=0, c1=None):
        c1 = self.ncols if c1 is None else c1
        k = c1 - c0
        self.ctx.ntt_batch_to(self.trace[c0], self.polys[c0], FP, self.log_n, k, inverse=True)
        if self.lde_fn is not None:
            self.lde_fn(c0, k)
        else:
            self.ctx.lde_batch(self.polys[c0], self.lde[c0], FP, self.log_n, self.log_b, k, offset=GENERATOR, bitrev=True)

    def commit(self):
        return self.ctx.merkle_commit(self.lde, FP, self.N, self.ncols, leaves=self.leaves, nodes=self.nodes)

    def evaluate(self):
        if self.evaluator is not None:
            self.evaluator.run(self.lde, self.ce)

    def run_resident(self):
        """whole step on a trace that is already in self.trace; returns the 32-byte root"""
        self.transform()
        root = self.commit()
        self.evaluate()
        return root

    # ---- end to end from host memory
    def run_from_host(self, host_trace, commit_fn=None, evaluate_fn=None):
        """host_trace: pinned (ncols, n) int64 tensor.  Returns (root, pinned host tensor of the composition
        evaluations or None).  H2D of column chunk k+1 overlaps iNTT + LDE of chunk k."""
        assert host_trace.is_pinned() and tuple(host_trace.shape) == (self.ncols, self.n)
        self.copy.wait_stream(self.compute)          # previous step must be done with self.trace
        events = []
        with torch.cuda.stream(self.copy):
            for c0 in range(0, self.ncols, self.chunk):
                c1 = min(c0 + self.chunk, self.ncols)
                self.trace[c0:c1].copy_(host_trace[c0:c1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy)
                events.append((c0, c1, ev))
        for c0, c1, ev in events:
            self.compute.wait_event(ev)
            self.transform(c0, c1)
        # (multi-GPU: commit_fn / evaluate_fn are the sharded versions, ministark_b200/parallel.py)
        root = (commit_fn or self.commit)()           # synchronises: the root is read back
        (evaluate_fn or self.evaluate)()
        if self.ce is not None:
            with torch.cuda.stream(self.compute):
                self.host_ce.copy_(self.ce, non_blocking=True)
        return root, self.host_ce
