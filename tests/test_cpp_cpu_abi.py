"""CPU-only: the C++ host layer run end to end WITHOUT a GPU.

oracle/libms_cpu_abi.so (oracle/cpu_abi.c) exports the entry points of include/ministark_b200.h on top of the CPU
oracle; it is test infrastructure, a different library from the product's, and nothing under ministark_b200/ loads it.
Linked against it,
  * include/ministark_prover.hpp (the C++ `default_prove`, tests/cpp/prover_test.cpp) must emit exactly the bytes of
    oracle/stark_oracle.cpu_prove — the restated reference prover — for examples/fib and examples/brainfuck (the latter
    through the C++ twin of the device extension-column builder), and its own C++ verifier must accept them;
  * include/ministark_gpu.hpp (the mirror of the Rust item set GpuFft / GpuIfft / Matrix / MatrixMerkleTree,
    tests/cpp/harness.cpp) must reproduce the oracle's polynomials and Merkle root;
  * the flattened evaluator programs of ministark_b200/expr.py::compile_program, executed by the compiled chunked
    interpreter behind ms_eval_constraints, must equal the tree-walking oracle (oracle/eval_oracle.py) — the same
    cases as the GPU kernel's parity test (tests/test_gpu_eval.py).
The GPU runs of the same binaries are tests/test_zz_gpu_cpp_prover.py and tests/test_gpu_cpp_mirror.py."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from ministark_b200 import expr as E
from ministark_b200.air import Air, ProofOptions
from ministark_b200.examples import fib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
P = E.P
GENERATOR = 7 * 2**64 % P


@pytest.fixture(scope="module")
def cpu_abi(orc):
    subprocess.check_call(["make", "-s", "-C", ORACLE, "libms_cpu_abi.so"])
    return os.path.join(ORACLE, "libms_cpu_abi.so")


def _build(tmp, cpu_abi, source):
    exe = tmp / source.replace(".cpp", "_cpu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", source),
                           "-o", str(exe), "-L", ORACLE, "-lms_cpu_abi", f"-Wl,-rpath,{ORACLE}"])
    return exe


@pytest.fixture(scope="module")
def prover_exe(tmp_path_factory, cpu_abi):
    return _build(tmp_path_factory.mktemp("cpp_cpu"), cpu_abi, "prover_test.cpp")


@pytest.mark.parametrize("log_rows,opts", [(7, (16, 4, 4, 8, 16)), (6, (10, 2, 0, 2, 8)), (10, (32, 4, 8, 8, 64)), (8, (16, 16, 3, 16, 4))])
def test_cpp_prover_fib_bytes_equal_cpu_restatement(prover_exe, orc, log_rows, opts):
    from oracle import stark_oracle as SO
    out = subprocess.run([str(prover_exe), "fib", str(log_rows)] + [str(o) for o in opts], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    last, hexbytes = out.stdout.split()
    trace, want_last = fib.gen_trace(8 << log_rows)
    assert int(last) == want_last
    claim = fib.FibClaim(want_last)
    mk = lambda n, o: Air(claim.AirConfig, n, claim.get_public_inputs(), ProofOptions(*o))
    want = SO.cpu_prove(claim, opts, trace.base_columns(), mk)
    assert bytes.fromhex(hexbytes) == want
    SO.verify(claim, bytes.fromhex(hexbytes), 10, mk)


@pytest.mark.parametrize("which,opts", [("hello", (19, 16, 20, 16, 16)), ("burner:4:4:4", (16, 16, 6, 8, 8))])
def test_cpp_prover_brainfuck_bytes_equal_cpu_restatement(prover_exe, orc, which, opts):
    """17 Fp + 9 Fq3 columns, ce_blowup 16: the extension columns come from bf::device_extension (fused evaluator +
    ms_scan_affine), the reference formulation builds them row by row (examples/brainfuck/trace.rs:108-279)"""
    from oracle import stark_oracle as SO
    from ministark_b200.examples import brainfuck as bf
    ii, mi = bf.test_rng_fq3(2)
    out = subprocess.run([str(prover_exe), "bf", which] + [str(o) for o in opts] + [str(v) for v in ii + mi], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr
    output_hex, hexbytes = out.stdout.split()
    src = bf.HELLO_WORLD if which == "hello" else bf.cycle_burner(*[int(v) for v in which.split(":")[1:]])
    trace, output = bf.simulate(src)
    assert bytes.fromhex(output_hex[len("out:"):]) == output
    claim = bf.BrainfuckClaim(src, b"", output)
    mk = lambda n, o: Air(claim.AirConfig, n, claim, ProofOptions(*o))
    want = SO.cpu_prove(claim, opts, trace.base_columns(), mk, ext_builder=trace.build_extension_columns)
    assert bytes.fromhex(hexbytes) == want


def test_bench_baseline_binary(cpu_abi, orc):
    """oracle/cpu_prover — what `bench.py --impl reference` times as the compiled CPU prove — emits the same bytes"""
    import json
    from oracle import stark_oracle as SO
    from ministark_b200.examples import brainfuck as bf
    subprocess.check_call(["make", "-s", "-C", ORACLE, "cpu_prover"])
    exe = os.path.join(ORACLE, "cpu_prover")
    opts = (32, 4, 8, 8, 64)
    r = json.loads(subprocess.run([exe, "8"] + [str(o) for o in opts] + ["--hex"], capture_output=True, text=True, check=True).stdout)
    trace, last = fib.gen_trace(8 << 8)
    claim = fib.FibClaim(last)
    mk = lambda n, o: Air(claim.AirConfig, n, claim.get_public_inputs(), ProofOptions(*o))
    assert r["verified"] and r["rows"] == 256 and bytes.fromhex(r["proof_hex"]) == SO.cpu_prove(claim, opts, trace.base_columns(), mk)
    ii, mi = bf.test_rng_fq3(2)
    opts = (16, 16, 6, 8, 8)
    r = json.loads(subprocess.run([exe, "bf", "4", "4", "4"] + [str(o) for o in opts] + [str(v) for v in ii + mi] + ["--hex"], capture_output=True,
                                  text=True, check=True).stdout)
    src = bf.cycle_burner(4, 4, 4)
    trace, output = bf.simulate(src)
    claim = bf.BrainfuckClaim(src, b"", output)
    mk = lambda n, o: Air(claim.AirConfig, n, claim, ProofOptions(*o))
    assert r["rows"] == 1024 and bytes.fromhex(r["proof_hex"]) == SO.cpu_prove(claim, opts, trace.base_columns(), mk,
                                                                                ext_builder=trace.build_extension_columns)


def test_cpp_prover_reports_library_errors(prover_exe):
    """the reference panics (gpu/src/stage.rs:55-75); the C++ layer turns a non-zero status into an exception carrying
    ms_last_error(): a 1-row FRI layer cannot be committed (merkle.rs:113-128 needs >= 2 leaves), and options whose
    folding factor exceeds the domain are refused before any size underflows"""
    out = subprocess.run([str(prover_exe), "fib", "1", "4", "2", "0", "4", "1"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "fri layer commit: merkle tree needs a power-of-two number of leaves >= 2, got 1" in out.stderr
    out = subprocess.run([str(prover_exe), "fib", "1", "4", "2", "0", "16", "1"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "smaller than the folding factor" in out.stderr


@pytest.mark.parametrize("log_n,ncols,log_b", [(8, 3, 2), (11, 5, 3), (4, 1, 1)])
def test_cpp_mirror_of_the_rust_item_set(tmp_path_factory, cpu_abi, orc, log_n, ncols, log_b):
    exe = _build(tmp_path_factory.mktemp("cpp_cpu_h"), cpu_abi, "harness.cpp")
    tmp = tmp_path_factory.mktemp("io")
    trace = orc.rand_matrix(ncols, 1 << log_n, 1, seed=log_n)
    trace.tofile(tmp / "in.bin")
    out = subprocess.run([str(exe), str(tmp / "in.bin"), str(tmp / "out.bin"), str(log_n), str(ncols), str(log_b)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    blob = (tmp / "out.bin").read_bytes()
    polys = orc.ntt(trace, 1, log_n, inverse=True)
    lde = orc.lde(polys, 1, log_n, log_b, orc.generator(), True)
    assert blob[:32] == orc.merkle_nodes(orc.hash_rows(lde, 1))[1].tobytes()
    assert np.array_equal(np.frombuffer(blob[32:], dtype=np.uint64).reshape(ncols, -1), polys)


def test_cpu_abi_covers_the_header_except_the_device_only_entry_points(cpu_abi):
    """every symbol of include/ministark_b200.h is either exported by the CPU build or on the short list of entry points
    with no CPU meaning (peer memory, the NVRTC check, the lazy-arithmetic probe) — a new header symbol shows up here"""
    from ministark_b200 import _lib
    lib = C.CDLL(cpu_abi)
    device_only = {"ms_ipc_export", "ms_ipc_open", "ms_ipc_close", "ms_lde_batch_scatter", "ms_eval_jit_check", "ms_debug_lazy_ops"}
    missing = {s for s in _lib.header_symbols() if not hasattr(lib, s)}
    assert missing == device_only
    lib.ms_version.restype = C.c_char_p
    assert b"CPU oracle" in lib.ms_version() and b"sm_100a" not in lib.ms_version()     # cannot be mistaken for the product


# ---------------------------------------------------------------------------------------------- evaluator programs
class _Abi:
    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.lib.ms_last_error.restype = C.c_char_p
        self.h = C.c_void_p()
        assert self.lib.ms_ctx_create(0, C.byref(self.h)) == 0

    def eval(self, prog, log_m, base, ext, fq, bitrev=False, out_bitrev=False, extra=()):
        out = np.zeros((1 << log_m) * fq, dtype=np.uint64)
        cols = ([] if base is None else [np.ascontiguousarray(c) for c in base]) + ([] if ext is None else [np.ascontiguousarray(c) for c in ext])
        isq = [0] * (0 if base is None else len(base)) + [1] * (0 if ext is None else len(ext))
        for arr, q in extra:
            cols.append(np.ascontiguousarray(arr))
            isq.append(int(q))
        ptrs = (C.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])
        flags = (C.c_int * max(len(cols), 1))(*isq)
        code, consts = np.ascontiguousarray(prog.code), np.ascontiguousarray(prog.consts)
        rc = self.lib.ms_eval_constraints_ptrs(self.h, C.c_void_p(code.ctypes.data), len(prog), C.c_void_p(consts.ctypes.data), consts.shape[0],
                                               ptrs, flags, len(cols), fq, log_m, C.c_uint64(GENERATOR), int(bitrev), int(out_bitrev),
                                               C.c_void_p(out.ctypes.data))
        if rc:
            raise ValueError(self.lib.ms_last_error(self.h).decode())
        return out


@pytest.fixture(scope="module")
def abi(cpu_abi):
    return _Abi(cpu_abi)


@pytest.mark.parametrize("bitrev", [False, True])
@pytest.mark.parametrize("blowup", [1, 4])
def test_compiled_programs_match_the_tree_oracle(abi, orc, bitrev, blowup):
    from oracle import eval_oracle
    log_m = 9 + (blowup.bit_length() - 1)
    m = 1 << log_m
    rng = random.Random(7)
    base = orc.rand_matrix(3, m, 1, seed=1)
    ext = orc.rand_matrix(2, m, 3, seed=2)
    base[0][5] = 0                                   # a zero under a division: batch inversion must leave it zero
    chal = [tuple(rng.randrange(P) for _ in range(3)) for _ in range(2)]
    hint = [tuple(rng.randrange(P) for _ in range(3))]
    x = E.X()
    cases = [
        x,
        x * x + 5,
        (x ** 3 - 1) / (x - 3),
        E.Trace(0, 0) * E.Trace(1, 1) - E.Trace(2, -1),
        E.Trace(3, 0) * E.Trace(0, 1) + E.Trace(4, 2) * E.Trace(3, -2),
        (E.Trace(3, 1) + E.Challenge(0)) / (E.Trace(4, 0) - E.Hint(0)),
        E.Constant((1, 2, 3)) * x ** 5 - E.Challenge(1) ** 3,
        -(E.Trace(1, 0) ** 7) + E.Constant(9),
        E.Trace(1, 0) / E.Trace(0, 0),
        E.Constant(4) * E.Constant(5) + E.Challenge(0) / E.Challenge(1),
    ]
    b = np.stack([orc.bit_reverse(c, 1, log_m) for c in base]) if bitrev else base
    e = np.stack([orc.bit_reverse(c, 3, log_m) for c in ext]) if bitrev else ext
    for ex in cases:
        prog = E.compile_program(ex, 3, chal, hint, lde_step=blowup, log_ce=log_m)
        want = eval_oracle.evaluate(ex.to_tuple(), log_m, orc.generator(), base, ext, fq_lanes=3, challenges=chal, hints=hint, lde_step=blowup)
        assert np.array_equal(abi.eval(prog, log_m, b, e, 3, bitrev=bitrev), want)
        if bitrev:      # out_bitrev: the result stays in the storage order of the inputs
            assert np.array_equal(abi.eval(prog, log_m, b, e, 3, bitrev=True, out_bitrev=True), orc.bit_reverse(want, 3, log_m))


def test_fq_equals_fp_and_program_validation(abi, orc):
    from oracle import eval_oracle
    log_m = 10
    base = orc.rand_matrix(2, 1 << log_m, 1, seed=5)
    ex = (E.Trace(0, 1) - E.Trace(0, 0) * E.Trace(1, 0)) * (E.X() ** 2 * E.Challenge(0) + E.Challenge(1))
    prog = E.compile_program(ex, 2, [11, 12], (), log_ce=log_m)
    want = eval_oracle.evaluate(ex.to_tuple(), log_m, orc.generator(), base, None, fq_lanes=1, challenges=[11, 12], hints=(), lde_step=1)
    assert np.array_equal(abi.eval(prog, log_m, base, None, 1), want)
    with pytest.raises(ValueError, match="out of range"):                 # column 5 referenced, 2 provided (eval_cpu.rs:148 panics)
        abi.eval(E.compile_program(E.Trace(5, 0) + 1, 8, log_ce=4), 4, base[:, :16], None, 1)
    bad = E.compile_program(E.Trace(0, 0) + 1, 2, log_ce=4)
    bad.code = bad.code.copy()
    bad.code[-1, 2] = 40                                                  # STORE of a register nothing wrote
    with pytest.raises(ValueError, match="before it is written"):
        abi.eval(bad, 4, base[:, :16], None, 1)


@pytest.mark.parametrize("blowup", [1, 4])
def test_periodic_columns_through_the_compiled_interpreter(abi, orc, blowup):
    """Periodic(coeffs, interval) leaves (src/constraints.rs:107-146, src/eval_cpu.rs:234-256); the table is the oracle's
    (the product builds it on the device: expr.periodic_tables)"""
    from oracle import eval_oracle
    log_n = 7
    log_m = log_n + (blowup.bit_length() - 1)
    m = 1 << log_m
    base = orc.rand_matrix(2, m, 1, seed=21)
    pa = E.Periodic([3, 5, 11, 2], 8)
    ex = E.Trace(0, 0) * pa - E.Trace(1, 1) + pa * pa * E.X()
    prog = E.compile_program(ex, 2, lde_step=blowup, log_ce=log_m, num_cols=2)
    assert len(prog.periodic) == 1
    want = eval_oracle.evaluate(ex.to_tuple(), log_m, orc.generator(), base, None, fq_lanes=1, lde_step=blowup)
    # the table by its definition (big-integer spec): P(y) at y = (offset * g_m^i)^(n / interval), i < interval * lde_step
    from oracle import pyspec as S
    tabs = []
    for slot, coeffs, interval, is_q, log_len in prog.periodic:
        assert slot == 2 and not is_q and (1 << log_len) == interval * blowup
        g, n = S.root_of_unity(log_m), 1 << log_n
        ys = [pow(7 * pow(g, i, S.P) % S.P, n // interval, S.P) for i in range(interval * blowup)]
        tabs.append(np.array([S.to_mont(sum(c * pow(y, k, S.P) for k, c in enumerate(coeffs)) % S.P) for y in ys], dtype=np.uint64))
    got = abi.eval(prog, log_m, base, None, 1, extra=[(t, False) for t in tabs])
    assert np.array_equal(got, want)
