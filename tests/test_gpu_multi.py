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
        sc = parallel.ShardedCommit(parallel.CudaEngine(ctx, torch.device("cuda", rank)), dist, log_n, log_b, ncols, fused=fused)
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
