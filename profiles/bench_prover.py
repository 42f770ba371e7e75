#!/usr/bin/env python
"""End-to-end prover timing (SURVEY.md §8d config 5 substitute: examples/fib, ProofOptions(32, 4, 8, 8, 64)).
One JSON line per size: wall-clock seconds of GpuProver.prove (host trace -> proof object, incl. H2D, all commitments,
FRI, PoW, queries) with the per-phase split, proof size, and — for --cpu-log-rows — the reference-formulation CPU
prover (oracle/stark_oracle.cpu_prove) on this box's host cores with the verifier's verdict on the GPU proof."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ministark_b200.air import Air, ProofOptions
from ministark_b200.examples import fib
from ministark_b200.prover import GpuProver

ap = argparse.ArgumentParser()
ap.add_argument("--log-rows", type=int, nargs="+", default=[18, 21, 22])
ap.add_argument("--cpu-log-rows", type=int, nargs="*", default=[18])
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
prover = GpuProver(0)
for lr in args.log_rows:
    trace, last = fib.gen_trace(8 << lr, pinned=True)            # page-locked columns (the reference's GpuAllocator role)
    pageable, _ = fib.gen_trace(8 << lr)
    claim = fib.FibClaim(last)
    prover.prove(claim, fib.OPTIONS, trace)                      # warm-up: plans, twiddle tables, NVRTC

    def best_of(tr):
        best = None
        for _ in range(args.reps):
            t = time.perf_counter()
            proof = prover.prove(claim, fib.OPTIONS, tr)
            dt = time.perf_counter() - t
            if best is None or dt < best[0]:
                best = (dt, proof)
        return best

    dt, proof = best_of(trace)
    dt_pageable, _ = best_of(pageable)
    trace = pageable
    line = {"bench": "fib_prove", "log_rows": lr, "cols": 8, "options": [32, 4, 8, 8, 64], "gpu_prove_s": dt,
            "gpu_prove_pageable_trace_s": dt_pageable,
            "phases_s": {k: round(v, 5) for k, v in proof.timings.items()}, "proof_bytes": len(proof.to_bytes())}
    if lr in args.cpu_log_rows:
        from oracle import stark_oracle as SO
        from oracle import oracle as orc
        mk = lambda n, o: Air(claim.AirConfig, n, claim.get_public_inputs(), ProofOptions(*o))
        tm = {}
        t = time.perf_counter()
        want = SO.cpu_prove(claim, (32, 4, 8, 8, 64), trace.base_columns(), mk, timings=tm)
        line["cpu_prove_s"] = time.perf_counter() - t
        line["cpu_threads"] = orc.num_threads()
        line["cpu_phases_s"] = {k: round(v, 4) for k, v in tm.items()}
        line["bytes_identical"] = want == proof.to_bytes()
        t = time.perf_counter()
        SO.verify(claim, proof.to_bytes(), fib.SECURITY_LEVEL, mk)
        line["verify_s"] = time.perf_counter() - t
        line["verified"] = True
    print(json.dumps(line), flush=True)
