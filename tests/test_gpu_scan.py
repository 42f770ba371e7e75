"""GPU: ms_scan_affine (running products / running evaluations as a parallel scan, csrc/scan.cu) vs the oracle's
sequential loop — the form the reference uses to build extension columns (examples/brainfuck/trace.rs:108-279)."""
import numpy as np
import pytest

import ministark_b200 as ms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return ms.Context(0)


@pytest.mark.parametrize("n", [1, 2, 7, 255, 2048, 2049, 5000, (1 << 16) + 3, 1 << 21])
@pytest.mark.parametrize("field", [ms.FP, ms.FQ3])
def test_scan_affine_matches_sequential_loop(ctx, orc, n, field):
    rng = np.random.default_rng(n * 7 + field)
    init = orc.rand_matrix(1, 1, 3, seed=n)[0].copy()
    if field == 1:
        init[1:] = 0
    cases = [
        dict(a=orc.rand_matrix(1, n, field, seed=11)[0], fa=field, inclusive=False),                       # running product
        dict(a=orc.rand_matrix(1, n, 1, seed=12)[0], fa=1, b=orc.rand_matrix(1, n, 1, seed=13)[0], fb=1, inclusive=True),
        dict(a_const=orc.rand_matrix(1, 1, 3, seed=14)[0] * np.array([1] + [int(field == 3)] * 2, dtype=np.uint64),
             b=orc.rand_matrix(1, n, 1, seed=15)[0], fb=1, inclusive=True),                                # running evaluation
        dict(a=orc.rand_matrix(1, n, field, seed=16)[0], fa=field, b=orc.rand_matrix(1, n, field, seed=17)[0], fb=field,
             inclusive=False),
    ]
    for k, cs in enumerate(cases):
        if n > (1 << 16) + 3 and k in (1, 3):
            continue
        a, b = cs.get("a"), cs.get("b")
        if a is not None and n > 4:                 # masked rows (a = 1, b = 0) and zeros, as padding rows produce
            a = a.copy()
            a.reshape(n, cs["fa"])[rng.integers(0, n, size=3)] = np.array([ms.ONE] + [0] * (cs["fa"] - 1), dtype=np.uint64)
            a.reshape(n, cs["fa"])[rng.integers(0, n)] = 0
        want = orc.scan_affine(field, n, init, a=a, fa=cs.get("fa", 1), a_const=cs.get("a_const"), b=b, fb=cs.get("fb", 1),
                               inclusive=cs["inclusive"])
        got = np.empty(n * field, dtype=np.uint64)
        ctx.scan_affine(got, field, n, init[:field], a=a, a_field=cs.get("fa", 1), a_const=cs.get("a_const"),
                        b=b, b_field=cs.get("fb", 1), inclusive=cs["inclusive"])
        assert np.array_equal(got, want), (n, field, k)


def test_scan_affine_resident_and_errors(ctx, orc):
    torch = pytest.importorskip("torch")
    n = 10000
    a = orc.rand_matrix(1, n, 3, seed=1)[0]
    init = np.array([ms.ONE, 0, 0], dtype=np.uint64)
    d_a = torch.from_numpy(a.view(np.int64)).cuda()
    d_out = torch.empty(3 * n, dtype=torch.int64, device="cuda")
    ctx.scan_affine(d_out, ms.FQ3, n, init, a=d_a, a_field=ms.FQ3)
    ctx.sync()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), orc.scan_affine(3, n, init, a=a, fa=3))
    with pytest.raises(ms.MsError):
        ctx.scan_affine(d_out, ms.FQ3, n, init)                                     # neither a nor a_const
    with pytest.raises(ms.MsError):
        ctx.scan_affine(d_out, ms.FP, n, init[:1], a=d_a, a_field=ms.FQ3)           # Fq3 multipliers into an Fp scan
