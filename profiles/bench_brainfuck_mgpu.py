#!/usr/bin/env python
"""examples/brainfuck at scale over the GPUs of one box (the north-star's "2^24-row brainfuck trace"):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        profiles/bench_brainfuck_mgpu.py 80 80 60          # cycle_burner(a, b, c): (80,80,60) pads to 2^22 rows

Every rank runs the VM (the trace is a deterministic function of the program), the sharded prover
(ministark_b200/prover_mgpu.py) proves it — 17 Fp + 9 Fq3 columns, ProofOptions(19, 16, 20, 16, 16): LDE domain 16 n,
16 Fq3 composition columns — and rank 0 checks the proof with the restated verifier.  One JSON line."""
import hashlib
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ministark_b200.air import Air, ProofOptions  # noqa: E402
from ministark_b200.examples import brainfuck as bf  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ministark_b200.prover_mgpu import ShardedProver
    prover = ShardedProver(dist, local)
else:
    from ministark_b200.prover import GpuProver
    prover = GpuProver(local)
a, b, c = (int(v) for v in sys.argv[1:4])
src = bf.cycle_burner(a, b, c)
t = time.perf_counter()
trace, out = bf.simulate(src)
t_sim = time.perf_counter() - t
claim = bf.BrainfuckClaim(src, b"", out)
times = []
proof = None
for _ in range(3):
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    proof = prover.prove(claim, bf.OPTIONS, trace)
    dt = torch.tensor([time.perf_counter() - t], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    times.append(float(dt.item()))
pb = proof.to_bytes()
digest = hashlib.sha256(pb).hexdigest()
if world > 1:
    all_d = [None] * world
    dist.all_gather_object(all_d, digest)
    assert len(set(all_d)) == 1, "ranks disagree on the proof"
if rank == 0:
    from oracle import stark_oracle as SO      # checker only
    t = time.perf_counter()
    SO.verify(claim, pb, bf.SECURITY_LEVEL, lambda n, o: Air(claim.AirConfig, n, claim, ProofOptions(*o)))
    free, total = torch.cuda.mem_get_info()
    print(json.dumps({"bench": "brainfuck_cycle_burner_multi_gpu", "program": f"cycle_burner({a},{b},{c})", "rows": len(trace),
                      "gpus": world, "cols": "17 Fp + 9 Fq3", "options": [19, 16, 20, 16, 16], "simulate_s_python_vm": t_sim,
                      "prove_s_runs": times, "prove_s": min(times[1:]),
                      "phases_s": {k: round(v, 5) for k, v in proof.timings.items()}, "proof_bytes": len(pb), "proof_sha256": digest,
                      "verify_s": time.perf_counter() - t, "verified": True,
                      "peak_device_memory_GiB_rank0": torch.cuda.max_memory_allocated() / 2**30}))
if world > 1:
    dist.destroy_process_group()
