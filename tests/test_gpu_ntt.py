"""GPU parity: NTT / iNTT / LDE through the C ABI vs the CPU oracle — bit exact.

Mirrors the reference's differential pattern (gpu/tests/shaders.rs:16-117: random polynomial,
GPU transform, assert_eq element-wise against ark-poly) with the oracle in arkworks' place, and
widens it: every size 2^0..2^21 (the reference tests 2048, 4096, 65536), plain and coset,
Fp and Fq3, host pointers (staged) and resident device pointers, batched columns."""
import numpy as np
import pytest

import ministark_b200 as ms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return ms.Context(0)


def _edge_column(n, lanes, rng):
    """random words seasoned with the values that break lazy/modular arithmetic"""
    P = ms.P
    edge = np.array([0, 1, P - 1, P - 2, 2**32 - 1, 2**32, 2**32 + 1, 0xFFFFFFFF00000000, 2**63, P - 2**32], dtype=np.uint64)
    v = rng.integers(0, P, size=n * lanes, dtype=np.uint64)
    k = min(len(edge), v.size)
    v[rng.choice(v.size, size=k, replace=False)] = edge[:k]
    return v


@pytest.mark.parametrize("log_n", list(range(0, 19)))
@pytest.mark.parametrize("coset", [False, True])
def test_fft_ifft_fp_all_sizes(ctx, orc, log_n, coset):
    rng = np.random.default_rng(log_n * 2 + coset)
    n = 1 << log_n
    offset = orc.generator() if coset else orc.ONE
    col = _edge_column(n, 1, rng)
    # GpuFft::encode / execute on a host slice (gpu/tests/shaders.rs:17-40)
    got = col.copy()
    fft = ms.GpuFft(ms.Domain(log_n, offset), ms.FP, ctx)
    fft.encode(got)
    fft.execute()
    want = orc.ntt(col.reshape(1, -1), 1, log_n, offset)[0]
    assert np.array_equal(got, want)
    # GpuIfft (gpu/tests/shaders.rs:94-117)
    ifft = ms.GpuIfft(ms.Domain(log_n, offset), ms.FP, ctx)
    back = got.copy()
    ifft.encode(back)
    ifft.execute()
    assert np.array_equal(back, col)
    assert np.array_equal(orc.ntt(want.reshape(1, -1), 1, log_n, offset, inverse=True)[0], col)


@pytest.mark.parametrize("log_n", [1, 4, 7, 11, 12, 13, 16])
@pytest.mark.parametrize("coset", [False, True])
def test_fft_fq3(ctx, orc, log_n, coset):
    # gpu/tests/shaders.rs:43-66: Fq3 coefficients, Fp twiddles
    rng = np.random.default_rng(100 + log_n)
    offset = orc.generator() if coset else orc.ONE
    col = _edge_column(1 << log_n, 3, rng)
    got = col.copy()
    fft = ms.GpuFft(ms.Domain(log_n, offset), ms.FQ3, ctx)
    fft.encode(got)
    fft.execute()
    want = orc.ntt(col.reshape(1, -1), 3, log_n, offset)[0]
    assert np.array_equal(got, want)
    ifft = ms.GpuIfft(ms.Domain(log_n, offset), ms.FQ3, ctx)
    ifft.encode(got)
    ifft.execute()
    assert np.array_equal(got, col)


@pytest.mark.parametrize("log_n,ncols", [(5, 7), (11, 17), (12, 9), (14, 5), (17, 3)])
def test_matrix_interpolate_multi_column(ctx, orc, log_n, ncols):
    # Matrix::interpolate encodes every column into one batch (src/matrix.rs:101-116)
    trace = orc.rand_matrix(ncols, 1 << log_n, 1, seed=log_n)
    m = ms.Matrix(trace, ms.FP, ctx)
    polys = m.interpolate(ms.Domain(log_n))
    assert np.array_equal(polys.cols, orc.ntt(trace, 1, log_n, inverse=True))
    # and many encodes on one plan, mixed lengths are rejected like the reference's assert_eq!
    fft = ms.GpuFft(ms.Domain(log_n), ms.FP, ctx)
    with pytest.raises(ms.MsError):
        fft.encode(np.zeros(3, dtype=np.uint64))


@pytest.mark.parametrize("log_n,log_b", [(0, 3), (2, 2), (3, 4), (4, 1), (5, 3), (9, 2), (11, 4), (12, 3), (13, 3), (16, 2), (17, 3)])
@pytest.mark.parametrize("bitrev", [True, False])
def test_lde_fp(ctx, orc, log_n, log_b, bitrev):
    # Matrix::(bit_reversed_)evaluate over the LDE coset offset = Fp::GENERATOR (src/matrix.rs:237-251)
    ncols = 3
    coeffs = orc.rand_matrix(ncols, 1 << log_n, 1, seed=7 * log_n + log_b)
    m = ms.Matrix(coeffs, ms.FP, ctx)
    dom = ms.Domain(log_n + log_b, ms.GENERATOR)
    got = (m.bit_reversed_evaluate(dom) if bitrev else m.evaluate(dom)).cols
    want = orc.lde(coeffs, 1, log_n, log_b, orc.generator(), bitrev=bitrev)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("log_n,log_b", [(4, 2), (11, 4), (13, 3)])
def test_lde_fq3(ctx, orc, log_n, log_b):
    coeffs = orc.rand_matrix(2, 1 << log_n, 3, seed=log_n)
    m = ms.Matrix(coeffs, ms.FQ3, ctx)
    got = m.bit_reversed_evaluate(ms.Domain(log_n + log_b, ms.GENERATOR)).cols
    assert np.array_equal(got, orc.lde(coeffs, 3, log_n, log_b, orc.generator(), bitrev=True))


def test_resident_device_pointers_and_strides(ctx, orc):
    torch = pytest.importorskip("torch")
    log_n, ncols, stride = 14, 4, (1 << 14) + 64   # padded column stride
    host = np.zeros((ncols, stride), dtype=np.uint64)
    trace = orc.rand_matrix(ncols, 1 << log_n, 1, seed=5)
    host[:, : 1 << log_n] = trace
    dev = torch.from_numpy(host.view(np.int64)).cuda()
    ctx.ntt_batch(dev, ms.FP, log_n, ncols, col_stride=stride, inverse=True)
    ctx.sync()
    got = dev.cpu().numpy().view(np.uint64)
    assert np.array_equal(got[:, : 1 << log_n], orc.ntt(trace, 1, log_n, inverse=True))
    assert not got[:, 1 << log_n:].any()      # padding untouched
    # resident LDE: coefficients -> evaluations, both on device
    out = torch.empty((ncols, 1 << (log_n + 2)), dtype=torch.int64, device="cuda")
    ctx.lde_batch(dev, out, ms.FP, log_n, 2, ncols, in_stride=stride)
    ctx.sync()
    want = orc.lde(orc.ntt(trace, 1, log_n, inverse=True), 1, log_n, 2, orc.generator(), bitrev=True)
    assert np.array_equal(out.cpu().numpy().view(np.uint64), want)


@pytest.mark.parametrize("log_n", [20])
def test_config1_roundtrip_2p20(ctx, orc, log_n):
    # BASELINE config 1: 2^20-point forward+inverse NTT, single column, bit-exact round trip
    col = orc.rand_matrix(1, 1 << log_n, 1, seed=1000)[0]
    for offset in (orc.ONE, orc.generator()):
        got = col.copy()
        ctx.ntt_batch(got, ms.FP, log_n, 1, offset=offset)
        assert np.array_equal(got, orc.ntt(col.reshape(1, -1), 1, log_n, offset)[0])
        ctx.ntt_batch(got, ms.FP, log_n, 1, inverse=True, offset=offset)
        assert np.array_equal(got, col)


def test_large_sizes_by_properties(ctx, orc):
    """2^24 (BASELINE config 3 column length): size-independent properties on device —
    round trip, linearity, and the LDE prefix property used by prover.rs:86-91."""
    torch = pytest.importorskip("torch")
    log_n = 24
    n = 1 << log_n
    a = torch.empty(n, dtype=torch.int64, device="cuda")
    b = torch.empty(n, dtype=torch.int64, device="cuda")
    ctx.fill_random(a, n, 1)
    ctx.fill_random(b, n, 2)
    s = torch.empty_like(a)
    ctx.pointwise("add", s, ms.FP, a, ms.FP, b, ms.FP, n=n)
    a0 = a.clone()
    for t in (a, b, s):
        ctx.ntt_batch(t, ms.FP, log_n, offset=ms.GENERATOR)
    s2 = torch.empty_like(a)
    ctx.pointwise("add", s2, ms.FP, a, ms.FP, b, ms.FP, n=n)
    ctx.sync()
    assert torch.equal(s, s2)                                  # NTT(a+b) == NTT(a)+NTT(b)
    ctx.ntt_batch(a, ms.FP, log_n, inverse=True, offset=ms.GENERATOR)
    ctx.sync()
    assert torch.equal(a, a0)                                  # iNTT(NTT(a)) == a
    # spot-check 64 evaluations against Horner on the host via the oracle's field ops
    coeffs = a0.cpu().numpy().view(np.uint64)
    ctx.ntt_batch(a, ms.FP, log_n, offset=ms.GENERATOR)
    ctx.sync()
    ev = a.cpu().numpy().view(np.uint64)
    g = orc.root_of_unity(log_n)
    for i in (1, n // 2 + 3, 9999999):
        x = orc.fp_mul(orc.generator(), orc.fp_pow(g, i))
        pt = np.array([x, 0, 0], dtype=np.uint64)
        assert int(orc.horner(coeffs, 1, pt)[0]) == int(ev[i])
