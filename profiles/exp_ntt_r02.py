"""Round-2 NTT experiment (run under gpurun): the persistent TMA pipeline (csrc/ntt_tma.cu) against the one-tile-per-CTA
passes (csrc/ntt.cu) — bit-exact equality first, then CUDA-event timings of the config-3 transforms.

  python profiles/exp_ntt_r02.py check      # small sizes vs the oracle, 2^24 vs the old path
  python profiles/exp_ntt_r02.py time       # 32 x 2^24: iNTT, LDE x8; old / TMA groups 2,3 / stage caps
Writes JSON lines to gpurun_out/exp_ntt_r02.jsonl."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ministark_b200 as ms  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "exp_ntt_r02.jsonl")


def emit(**kw):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(kw) + "\n")
    print(json.dumps(kw), flush=True)


def new_ctx():
    """a dedicated torch stream (the legacy default stream's handle is NULL, which ms_ctx_set_stream reads as "own stream")"""
    global _STREAM
    _STREAM = torch.cuda.Stream()
    torch.cuda.set_stream(_STREAM)
    return ms.Context(0, stream=_STREAM.cuda_stream)


def check():
    from oracle import oracle as orc
    ctx = new_ctx()
    ok = True
    # 2^16: passes [8, 8] -> strided (TMA) + contiguous (TMA) for the LDE; strided natural + transposing (old) for NTT
    for log_n, log_b, ncols in [(16, 3, 8), (16, 2, 16), (16, 0, 64)]:
        n = 1 << log_n
        coeffs = orc.rand_matrix(ncols, n, 1, seed=77 + log_b)
        coeffs[0, :8] = [0, 1, ms.P - 1, 2**32 - 1, 2**32, 2**63, ms.P - 2**32, 0xFFFFFFFF00000000]
        dev = torch.from_numpy(coeffs.view(np.int64)).cuda()
        out = torch.empty((ncols, n << log_b), dtype=torch.int64, device="cuda")
        res = {}
        for tma in (1, 0):
            ctx.set_option("ntt_tma", tma)
            out.zero_()
            l0 = ctx.launches
            ctx.lde_batch(dev, out, ms.FP, log_n, log_b, ncols)
            ctx.sync()
            res[tma] = (out.cpu().numpy().view(np.uint64).copy(), ctx.launches - l0)
        want = orc.lde(coeffs, 1, log_n, log_b, orc.generator(), bitrev=True)
        e1, e0 = bool(np.array_equal(res[1][0], want)), bool(np.array_equal(res[0][0], want))
        emit(test="lde_vs_oracle", log_n=log_n, log_b=log_b, ncols=ncols, tma_ok=e1, old_ok=e0)
        ok &= e1 and e0
    # natural-order forward / inverse NTT, 64 columns of 2^16 (strided pass through the TMA path, natural digit)
    for inverse in (False, True):
        for offset in (ms.ONE, ms.GENERATOR):
            ncols, log_n = 64, 16
            cols = orc.rand_matrix(ncols, 1 << log_n, 1, seed=5)
            want = orc.ntt(cols, 1, log_n, offset, inverse=inverse)
            for tma in (1, 0):
                ctx.set_option("ntt_tma", tma)
                d = torch.from_numpy(cols.view(np.int64)).cuda()
                ctx.ntt_batch(d, ms.FP, log_n, ncols, inverse=inverse, offset=offset)
                ctx.sync()
                e = bool(np.array_equal(d.cpu().numpy().view(np.uint64), want))
                emit(test="ntt_vs_oracle", inverse=inverse, coset=offset != ms.ONE, tma=tma, ok=e)
                ok &= e
    # 2^24: three passes; TMA (groups 2 and 3) against the old path, bit for bit
    log_n, log_b, ncols = 24, 3, 2
    n = 1 << log_n
    a = torch.empty((ncols, n), dtype=torch.int64, device="cuda")
    ctx.fill_random(a, ncols * n, 11)
    ref = torch.empty((ncols, n << log_b), dtype=torch.int64, device="cuda")
    ctx.set_option("ntt_tma", 0)
    ctx.lde_batch(a, ref, ms.FP, log_n, log_b, ncols)
    ctx.sync()
    got = torch.empty_like(ref)
    for groups in (2, 3):
        ctx.set_option("ntt_tma", 1)
        ctx.set_option("ntt_tma_groups", groups)
        got.zero_()
        ctx.lde_batch(a, got, ms.FP, log_n, log_b, ncols)
        ctx.sync()
        e = bool(torch.equal(got, ref))
        emit(test="lde_2p24_tma_vs_old", groups=groups, ok=e)
        ok &= e
    # iNTT / NTT 2^24 natural, 4 columns
    b = torch.empty((4, n), dtype=torch.int64, device="cuda")
    ctx.fill_random(b, 4 * n, 12)
    for inverse in (True, False):
        outs = {}
        for tma in (0, 1):
            ctx.set_option("ntt_tma", tma)
            ctx.set_option("ntt_tma_groups", 2)
            d = b.clone()
            ctx.ntt_batch(d, ms.FP, log_n, 4, inverse=inverse, offset=ms.GENERATOR)
            ctx.sync()
            outs[tma] = d
        e = bool(torch.equal(outs[0], outs[1]))
        emit(test="ntt_2p24_tma_vs_old", inverse=inverse, ok=e)
        ok &= e
    emit(test="check_all", ok=ok)
    return ok


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


def bench():
    ctx = new_ctx()
    log_n, log_b, ncols = 24, 3, 32
    n = 1 << log_n
    tr = torch.empty((ncols, n), dtype=torch.int64, device="cuda")
    ctx.fill_random(tr, ncols * n, 3)
    lde = torch.empty((ncols, n << log_b), dtype=torch.int64, device="cuda")
    configs = [("old", 0, 2, 8), ("tma_g2", 1, 2, 8), ("tma_g3", 1, 3, 8), ("tma_g2_s4", 1, 2, 4), ("tma_g2_s5", 1, 2, 5),
               ("tma_g3_s4", 1, 3, 4), ("tma_g2", 1, 2, 8), ("old", 0, 2, 8)]
    for name, tma, groups, stages in configs:
        ctx.set_option("ntt_tma", tma)
        ctx.set_option("ntt_tma_groups", groups)
        ctx.set_option("ntt_tma_stages", stages)
        mn, med = timeit(lambda: ctx.lde_batch(tr, lde, ms.FP, log_n, log_b, ncols))
        emit(bench="lde_32x2p24_x8", variant=name, ms_min=round(mn, 3), ms_med=round(med, 3))
        mn, med = timeit(lambda: ctx.ntt_batch(tr, ms.FP, log_n, ncols, inverse=True))
        emit(bench="intt_32x2p24", variant=name, ms_min=round(mn, 3), ms_med=round(med, 3))


def prof():
    """one small LDE (4 columns x 2^24, x8) for ncu: run under `ncu -k regex:ntt_tma -s 3 -c 3 --set full`"""
    ctx = new_ctx()
    log_n, log_b, ncols = 24, 3, 4
    n = 1 << log_n
    tr = torch.empty((ncols, n), dtype=torch.int64, device="cuda")
    ctx.fill_random(tr, ncols * n, 3)
    lde = torch.empty((ncols, n << log_b), dtype=torch.int64, device="cuda")
    for _ in range(2):
        ctx.lde_batch(tr, lde, ms.FP, log_n, log_b, ncols)
    ctx.sync()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "check"
    t0 = time.time()
    if what == "check":
        sys.exit(0 if check() else 1)
    if what == "prof":
        prof()
    else:
        bench()
    print("elapsed", time.time() - t0)
