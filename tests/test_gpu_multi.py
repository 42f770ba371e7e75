"""GPU, 2 ranks over NCCL (skipped on a 1-GPU box): column-sharded LDE -> all-to-all -> row-sharded Merkle
commit (ministark_b200/parallel.py) gives the same root as the oracle's single tree; the same host logic is
covered on CPU by tests/test_parallel_gloo.py."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, log_n, log_b, ncols, q, fused=None):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import ministark_b200 as ms
    from ministark_b200 import parallel
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        stream = torch.cuda.Stream(device=rank)
        torch.cuda.set_stream(stream)
        ctx = ms.Context(rank, stream=stream.cuda_stream)
        full = orc.rand_matrix(ncols, 1 << log_n, 1, seed=88)
        sc = parallel.ShardedCommit(parallel.CudaEngine(ctx, torch.device("cuda", rank), stream=stream), dist, log_n, log_b, ncols, fused=fused)
        assert sc.fused == (fused is not False)
        local = torch.from_numpy(full[sc.lo:sc.hi].view(np.int64).copy()).cuda(rank)
        roots = []
        for _ in range(2):               # twice: the second round overwrites slabs the peers have just hashed
            sc.transform(local)
            roots.append(sc.commit())
        assert roots[0] == roots[1]
        # the local copy of the ce-domain prefix (block 0) that the constraint evaluation of this rank reads
        ctx.sync()
        n = 1 << log_n
        prefix = sc.lde[:, :n].cpu().numpy().view(np.uint64)
        polys = orc.ntt(full[sc.lo:sc.hi], 1, log_n, inverse=True)
        assert np.array_equal(prefix, orc.lde(polys, 1, log_n, log_b, orc.generator(), True)[:, :n])
        q.put((rank, roots[0]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fused,log_n", [(None, 12), (False, 12), (None, 14), (None, 6)])
def test_sharded_commit_two_gpus(orc, fused, log_n):
    """fused: the LDE's last pass stores the coset blocks into the peers' row slabs (CUDA IPC over NVLink);
    not fused: LDE, then the NCCL all-to-all"""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    log_b, ncols = 3, 8
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, log_n, log_b, ncols, q, fused)) for r in range(2)]
    for p in procs:
        p.start()
    roots = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = orc.rand_matrix(ncols, 1 << log_n, 1, seed=88)
    lde = orc.lde(orc.ntt(full, 1, log_n, inverse=True), 1, log_n, log_b, orc.generator(), True)
    want = orc.merkle_nodes(orc.hash_rows(lde, 1))[1].tobytes()
    assert roots[0] == roots[1] == want


# ---- the whole prover over 2 GPUs (ministark_b200/prover_mgpu.py): proof bytes identical to the single-GPU prover's ----
def _prove_worker(rank, world, port, which, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from ministark_b200.air import Air, ProofOptions
    from ministark_b200.prover import GpuProver
    from ministark_b200.prover_mgpu import ShardedProver
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        claim, opts, trace = _make_case(which)
        sharded = ShardedProver(dist, rank).prove(claim, ProofOptions(*opts), trace).to_bytes()
        again = ShardedProver(dist, rank).prove(claim, ProofOptions(*opts), trace).to_bytes()
        single = GpuProver(rank).prove(claim, ProofOptions(*opts), trace).to_bytes() if rank == 0 else None
        q.put((rank, sharded, again, single))
    finally:
        dist.destroy_process_group()


def _make_case(which):
    from ministark_b200.examples import brainfuck as bf
    from ministark_b200.examples import fib, perm
    if which.startswith("fib"):
        log_rows = int(which.split(":")[1])
        trace, last = fib.gen_trace(8 << log_rows)
        return fib.FibClaim(last), (32, 4, 8, 8, 64), trace
    if which == "perm":
        trace = perm.gen_trace(1 << 10, seed=3)
        return perm.PermClaim(), (16, 8, 4, 4, 8), trace
    src = bf.HELLO_WORLD
    trace, output = bf.simulate(src)
    return bf.BrainfuckClaim(src, b"", output), (19, 16, 20, 16, 16), trace


@pytest.mark.parametrize("which", ["fib:7", "fib:13", "perm", "brainfuck"])
def test_sharded_prover_bytes_equal_single_gpu(which):
    """every matrix row-sharded by LDE coset blocks, FRI layers sharded by rows, paths assembled from their owners:
    the proof must not change by a byte, and the restated verifier accepts it"""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from ministark_b200.air import Air, ProofOptions
    from oracle import stark_oracle
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_prove_worker, args=(r, 2, port, which, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, (a, b, c)) for r, a, b, c in (q.get(timeout=600) for _ in range(2)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][0] == res[1][0] == res[0][1] == res[1][1], "ranks disagree on the proof"
    assert res[0][0] == res[0][2], "sharded proof differs from the single-GPU proof"
    claim, opts, _ = _make_case(which)
    mk = lambda n, o: Air(claim.AirConfig, n, claim.get_public_inputs(), ProofOptions(*o))
    stark_oracle.verify(claim, res[0][0], 10, mk)
