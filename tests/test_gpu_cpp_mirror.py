"""GPU: the C++ mirror of the reference's Rust interface (include/ministark_gpu.hpp) driven by a
compiled harness through the C ABI — Matrix::interpolate, bit_reversed_evaluate,
MatrixMerkleTree::from_matrix, GpuFft/GpuIfft encode/execute and their assertion behaviour."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_harness_matches_oracle(orc, tmp_path):
    exe = tmp_path / "harness"
    libdir = os.path.join(ROOT, "ministark_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "harness.cpp"), "-o", str(exe),
                           "-L", libdir, "-l:libministark_b200.so", f"-Wl,-rpath,{libdir}"])
    log_n, ncols, log_b = 12, 5, 2
    trace = orc.rand_matrix(ncols, 1 << log_n, 1, seed=11)
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    trace.tofile(inp)
    subprocess.check_call([str(exe), str(inp), str(out), str(log_n), str(ncols), str(log_b)])
    raw = open(out, "rb").read()
    root, polys = raw[:32], np.frombuffer(raw[32:], dtype=np.uint64).reshape(ncols, -1)
    want_polys = orc.ntt(trace, 1, log_n, inverse=True)
    assert np.array_equal(polys, want_polys)
    lde = orc.lde(want_polys, 1, log_n, log_b, orc.generator(), bitrev=True)
    assert root == orc.merkle_nodes(orc.hash_rows(lde, 1))[1].tobytes()
