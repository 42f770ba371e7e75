"""
oracle.py — ctypes binding of the C CPU oracle (oracle/gl_oracle.c).
TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(ministark_b200/) must never import this module.

All arrays are numpy uint64 holding Montgomery words exactly as the reference
keeps them in memory (canonical x*2^64 mod p).  Matrices are column-major:
shape (ncols, nrows*lanes), C-contiguous; lanes = 1 (Fp) or 3 (Fq3, c0,c1,c2).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

u64 = C.c_uint64
u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


def build(force=False):
    src = os.path.join(_HERE, "gl_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        L = _lib
        for name in ("orc_fp_one", "orc_generator"):
            getattr(L, name).restype = u64
        for name in ("orc_fp_from_canonical", "orc_fp_to_canonical", "orc_fp_inv1"):
            f = getattr(L, name); f.restype = u64; f.argtypes = [u64]
        for name in ("orc_fp_mul1", "orc_fp_add1", "orc_fp_sub1", "orc_fp_pow1"):
            f = getattr(L, name); f.restype = u64; f.argtypes = [u64, u64]
        L.orc_root_of_unity.restype = u64; L.orc_root_of_unity.argtypes = [C.c_uint]
        L.orc_vec_from_canonical.argtypes = [u64p, C.c_size_t]
        L.orc_vec_to_canonical.argtypes = [u64p, C.c_size_t]
        L.orc_fq3_mul1.argtypes = [u64p, u64p, u64p]
        L.orc_fq3_inv1.argtypes = [u64p, u64p]
        L.orc_fq3_pow1.argtypes = [u64p, u64, u64p]
        L.orc_bit_reverse.argtypes = [u64p, C.c_uint, C.c_uint]
        L.orc_ntt_columns.argtypes = [u64p, C.c_size_t, C.c_uint, C.c_uint, C.c_uint, u64, C.c_int]
        L.orc_lde_columns.argtypes = [u64p, C.c_size_t, u64p, C.c_size_t, C.c_uint, C.c_uint, C.c_uint,
                                      C.c_uint, u64, C.c_int]
        L.orc_sha256.argtypes = [u8p, C.c_size_t, u8p]
        L.orc_hash_rows.argtypes = [u64p, C.c_size_t, C.c_uint, C.c_uint, C.c_size_t, u8p]
        L.orc_merkle_nodes.argtypes = [u8p, C.c_size_t, u8p]
        L.orc_pointwise.argtypes = [C.c_int, C.c_uint, u64p, C.c_uint, u64p, C.c_uint, u64p, C.c_size_t,
                                    C.c_size_t, u64]
        L.orc_pointwise_const.argtypes = [C.c_int, C.c_uint, u64p, C.c_uint, u64p, C.c_uint, u64p, C.c_size_t]
        L.orc_sum_columns.argtypes = [u64p, C.c_size_t, C.c_uint, C.c_uint, C.c_size_t, u64p]
        L.orc_fri_apply_drp.argtypes = [u64p, C.c_uint, C.c_uint, C.c_uint, u64, u64p, u64p]
        L.orc_horner.argtypes = [u64p, C.c_uint, C.c_size_t, u64p, u64p]
        L.orc_divide_out_points.argtypes = [u64p, C.c_size_t, u64p, u64p, C.c_uint]
        L.orc_degree_adjust.argtypes = [u64p, C.c_size_t, u64p, u64p]
        L.orc_scan_affine.argtypes = [C.c_uint, u64p, C.c_uint, u64p, u64p, C.c_uint, C.c_size_t, u64p, C.c_int, u64p]
        L.orc_splitmix_fill.argtypes = [u64p, C.c_size_t, u64]
        L.orc_pow_grind.restype = u64
        L.orc_pow_grind.argtypes = [u8p, C.c_uint]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
    return _lib


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _p8(a):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u8p)


OP = dict(mul=0, add=1, convert=2, inv=3, exp=4, neg=5, mulpow=6, fill=7, sub=8)

ONE = 4294967295
P = 2**64 - 2**32 + 1


def to_mont(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    lib().orc_vec_from_canonical(_p(a.reshape(-1)), a.size)
    return a


def from_mont(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    lib().orc_vec_to_canonical(_p(a.reshape(-1)), a.size)
    return a


def root_of_unity(log_n):
    return int(lib().orc_root_of_unity(log_n))


def generator():
    return int(lib().orc_generator())


def fp_pow(a, e):
    return int(lib().orc_fp_pow1(a, e))


def fp_mul(a, b):
    return int(lib().orc_fp_mul1(a, b))


def fp_inv(a):
    return int(lib().orc_fp_inv1(a))


def rand_matrix(ncols, nrows, lanes=1, seed=1):
    """synthetic uniform field elements (SURVEY.md §8d splitmix64 generator)."""
    m = np.empty((ncols, nrows * lanes), dtype=np.uint64)
    for c in range(ncols):
        lib().orc_splitmix_fill(_p(m[c]), nrows * lanes, (0x9E3779B97F4A7C15 ^ (seed + c)) & (2**64 - 1))
    return m


def bit_reverse(col, lanes, log_n):
    out = np.ascontiguousarray(col).copy()
    lib().orc_bit_reverse(_p(out), lanes, log_n)
    return out


def ntt(mat, lanes, log_n, offset=ONE, inverse=False):
    """natural-order (i)NTT of every column; returns a new matrix."""
    out = np.ascontiguousarray(mat).copy()
    ncols = out.shape[0]
    lib().orc_ntt_columns(_p(out.reshape(-1)), out.shape[1], ncols, lanes, log_n, offset, int(inverse))
    return out


def lde(mat, lanes, log_n, log_blowup, offset, bitrev=True):
    mat = np.ascontiguousarray(mat)
    ncols = mat.shape[0]
    out = np.empty((ncols, (lanes << log_n) << log_blowup), dtype=np.uint64)
    lib().orc_lde_columns(_p(mat.reshape(-1)), mat.shape[1], _p(out.reshape(-1)), out.shape[1], ncols, lanes,
                          log_n, log_blowup, offset, int(bitrev))
    return out


def sha256(b):
    a = np.frombuffer(b, dtype=np.uint8).copy() if len(b) else np.zeros(1, dtype=np.uint8)
    out = np.empty(32, dtype=np.uint8)
    lib().orc_sha256(_p8(a), len(b), _p8(out))
    return out.tobytes()


def hash_rows(mat, lanes):
    mat = np.ascontiguousarray(mat)
    ncols, nrows = mat.shape[0], mat.shape[1] // lanes
    out = np.empty((nrows, 32), dtype=np.uint8)
    lib().orc_hash_rows(_p(mat.reshape(-1)), mat.shape[1], ncols, lanes, nrows, _p8(out.reshape(-1)))
    return out


def merkle_nodes(leaves):
    leaves = np.ascontiguousarray(leaves)
    n = leaves.shape[0]
    out = np.empty((n, 32), dtype=np.uint8)
    lib().orc_merkle_nodes(_p8(leaves.reshape(-1)), n, _p8(out.reshape(-1)))
    return out


def pointwise(op, lhs, lfield, rhs=None, rfield=1, shift=0, exponent=0, dfield=None):
    if dfield is None:
        dfield = max(lfield, rfield if rhs is not None else 1)
    n = lhs.size // lfield
    dst = np.empty(n * dfield, dtype=np.uint64)
    lib().orc_pointwise(OP[op], dfield, _p(dst), lfield, _p(np.ascontiguousarray(lhs)), rfield,
                        _p(np.ascontiguousarray(rhs)) if rhs is not None else None, n, shift, exponent)
    return dst


def pointwise_const(op, lhs, lfield, cst, cfield, n=None, dfield=None):
    if dfield is None:
        dfield = max(lfield, cfield)
    if n is None:
        n = lhs.size // lfield
    dst = np.empty(n * dfield, dtype=np.uint64)
    cst = np.ascontiguousarray(cst, dtype=np.uint64)
    lib().orc_pointwise_const(OP[op], dfield, _p(dst), lfield,
                              _p(np.ascontiguousarray(lhs)) if lhs is not None else None, cfield, _p(cst), n)
    return dst


def sum_columns(mat, lanes):
    mat = np.ascontiguousarray(mat)
    n = mat.shape[1] // lanes
    acc = np.empty(n * lanes, dtype=np.uint64)
    lib().orc_sum_columns(_p(mat.reshape(-1)), mat.shape[1], mat.shape[0], lanes, n, _p(acc))
    return acc


def fri_apply_drp(evals, lanes, log_n, log_ff, alpha, offset=ONE):
    evals = np.ascontiguousarray(evals)
    out = np.empty((lanes << log_n) >> log_ff, dtype=np.uint64)
    alpha = np.ascontiguousarray(alpha, dtype=np.uint64)
    lib().orc_fri_apply_drp(_p(evals), lanes, log_n, log_ff, offset, _p(alpha), _p(out))
    return out


def horner(coeffs, cf, point):
    coeffs = np.ascontiguousarray(coeffs)
    out = np.empty(3, dtype=np.uint64)
    lib().orc_horner(_p(coeffs), cf, coeffs.size // cf, _p(np.ascontiguousarray(point, dtype=np.uint64)), _p(out))
    return out


def divide_out_points(coeffs, zs, cs):
    out = np.ascontiguousarray(coeffs).copy()
    zs = np.ascontiguousarray(zs, dtype=np.uint64).reshape(-1)
    cs = np.ascontiguousarray(cs, dtype=np.uint64).reshape(-1)
    lib().orc_divide_out_points(_p(out), out.size // 3, _p(zs), _p(cs), zs.size // 3)
    return out


def horner_jobs(jobs):
    """jobs: [(coeffs array, coefficient lanes, point as 3 Montgomery words)], all of the same length; evaluated across the
    host threads (get_ood_evals, src/composer.rs:60-83).  Returns (njobs, 3) Montgomery words."""
    k = len(jobs)
    out = np.empty((k, 3), dtype=np.uint64)
    if not k:
        return out
    arrs = [np.ascontiguousarray(j[0]) for j in jobs]
    n = arrs[0].size // jobs[0][1]
    assert all(a.size // j[1] == n for a, j in zip(arrs, jobs))
    ptrs = (C.c_void_p * k)(*[a.ctypes.data for a in arrs])
    cf = (C.c_uint * k)(*[j[1] for j in jobs])
    pts = np.ascontiguousarray(np.concatenate([np.asarray(j[2], dtype=np.uint64).reshape(3) for j in jobs]))
    f = lib().orc_horner_jobs
    f.restype, f.argtypes = None, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint, C.c_void_p]
    f(ptrs, cf, n, pts.ctypes.data, k, out.ctypes.data)
    return out


def divide_out_points_columns(cols, zs_per_col, cs_per_col):
    """cols: (ncols, 3n) Fq3 coefficient columns, divided IN PLACE, each by its own list of points (Montgomery word
    triples) with its own coefficients, columns across the host threads (into_deep_poly, src/composer.rs:108-158)"""
    assert cols.dtype == np.uint64 and cols.flags["C_CONTIGUOUS"]
    ncols = cols.shape[0]
    counts = (C.c_uint * max(ncols, 1))(*[np.asarray(z).size // 3 for z in zs_per_col])
    flat = lambda per: np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.uint64).reshape(-1) for v in per] or [np.zeros(0, np.uint64)]))
    zs, cs = flat(zs_per_col), flat(cs_per_col)
    assert zs.size == cs.size == 3 * sum(counts) and len(zs_per_col) == len(cs_per_col) == ncols
    f = lib().orc_divide_out_points_columns
    f.restype, f.argtypes = None, [C.c_void_p, C.c_size_t, C.c_uint, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    f(cols.ctypes.data, cols.shape[1], ncols, cols.shape[1] // 3, zs.ctypes.data, cs.ctypes.data, counts)
    return cols


def degree_adjust(coeffs, alpha, beta):
    out = np.ascontiguousarray(coeffs).copy()
    lib().orc_degree_adjust(_p(out), out.size // 3, _p(np.ascontiguousarray(alpha, dtype=np.uint64)),
                            _p(np.ascontiguousarray(beta, dtype=np.uint64)))
    return out


def scan_affine(field, n, init, a=None, fa=1, a_const=None, b=None, fb=1, inclusive=False):
    """x_0 = init, x_(i+1) = x_i * a_i + b_i; returns out[i] = x_i (or x_(i+1) if inclusive), n * field words"""
    out = np.empty(n * field, dtype=np.uint64)
    w3 = lambda v: np.ascontiguousarray(v, dtype=np.uint64)
    lib().orc_scan_affine(field, _p(np.ascontiguousarray(a)) if a is not None else None, fa,
                          _p(w3(a_const)) if a_const is not None else None,
                          _p(np.ascontiguousarray(b)) if b is not None else None, fb, n, _p(w3(init)), int(inclusive), _p(out))
    return out


def pow_grind(seed, bits):
    a = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
    return int(lib().orc_pow_grind(_p8(a), bits))


def num_threads():
    return int(lib().orc_num_threads())
