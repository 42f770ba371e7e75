#!/usr/bin/env python
"""experiment: Merkle commit time vs number of SHA-256 additions forced onto the FMA pipe (MS_SHA_FMA_ADDS)"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ministark_b200 as ms
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = ms.Context(0, stream=stream.cuda_stream)
log_rows, ncols = 26, 32
N = 1 << log_rows
lde = torch.empty((ncols, N), dtype=torch.int64, device="cuda")
ctx.fill_random(lde, ncols * N, 5)
leaves = torch.empty((N, 4), dtype=torch.int64, device="cuda"); nodes = torch.empty_like(leaves)
def run(): return ctx.merkle_commit(lde, ms.FP, N, ncols, leaves=leaves, nodes=nodes)
for _ in range(2): root = run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): run()
b.record(); torch.cuda.synchronize()
print(json.dumps({"variant": os.environ.get("MS_SHA_FMA_ADDS", "0"), "merkle_commit_ms_2p26x32": a.elapsed_time(b) / 5, "root": root.hex()[:16]}))
