"""ctypes binding of libministark_b200.so (include/ministark_b200.h).

The library is the product: if it is missing or no CUDA device is present, everything here
fails loudly — there is no CPU fallback and nothing from oracle/ is ever imported.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MS_LIB_PATH") or os.path.join(_HERE, "libministark_b200.so")   # MS_LIB_PATH: A/B builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ministark_b200.h")

u64 = C.c_uint64
vp = C.c_void_p
sz = C.c_size_t
ui = C.c_uint
ci = C.c_int

_SIGS = {
    "ms_ctx_create": (ci, [ci, C.POINTER(vp)]),
    "ms_ctx_destroy": (ci, [vp]),
    "ms_ctx_set_stream": (ci, [vp, vp]),
    "ms_ctx_sync": (ci, [vp]),
    "ms_last_error": (C.c_char_p, [vp]),
    "ms_version": (C.c_char_p, []),
    "ms_launch_count": (u64, [vp]),
    "ms_set_option": (ci, [vp, C.c_char_p, C.c_int64]),
    "ms_alloc_device": (ci, [vp, sz, C.POINTER(vp)]),
    "ms_alloc_host_pinned": (ci, [vp, sz, C.POINTER(vp)]),
    "ms_free": (ci, [vp, vp]),
    "ms_copy": (ci, [vp, vp, vp, sz]),
    "ms_ntt_plan_create": (ci, [vp, ci, ui, ci, u64, C.POINTER(vp)]),
    "ms_ntt_encode": (ci, [vp, vp]),
    "ms_ntt_execute": (ci, [vp]),
    "ms_ntt_plan_destroy": (ci, [vp]),
    "ms_ntt_batch": (ci, [vp, ci, vp, sz, ui, ui, ci, u64]),
    "ms_ntt_batch_to": (ci, [vp, ci, vp, sz, vp, sz, ui, ui, ci, u64]),
    "ms_lde_batch": (ci, [vp, ci, vp, sz, vp, sz, ui, ui, ui, u64, ci]),
    "ms_bit_reverse": (ci, [vp, ci, vp, sz, ui, ui]),
    "ms_pointwise": (ci, [vp, ci, ci, vp, ci, vp, ci, vp, sz, sz, u64]),
    "ms_pointwise_const": (ci, [vp, ci, ci, vp, ci, vp, ci, vp, sz]),
    "ms_sum_columns": (ci, [vp, ci, vp, sz, ui, sz, vp]),
    "ms_hash_rows_sha256": (ci, [vp, ci, vp, sz, ui, sz, vp]),
    "ms_merkle_nodes_sha256": (ci, [vp, vp, sz, vp]),
    "ms_merkle_commit_sha256": (ci, [vp, ci, vp, sz, ui, sz, vp, vp, vp]),
    "ms_merkle_commit_rows_sha256": (ci, [vp, vp, ui, sz, vp, vp, vp]),
    "ms_pow_grind_sha256": (ci, [vp, vp, ui, C.POINTER(u64)]),
    "ms_merkle_prove_sha256": (ci, [vp, vp, vp, sz, vp, ui, vp, vp, vp, vp]),
    "ms_matrix_from_rows": (ci, [vp, ci, vp, sz, ui, vp, sz]),
    "ms_gather_rows": (ci, [vp, ci, vp, sz, ui, sz, vp, ui, vp]),
    "ms_gather_rows_rowmajor": (ci, [vp, vp, ui, sz, vp, ui, vp]),
    "ms_debug_lazy_ops": (ci, [vp, vp, vp, sz, vp]),
    "ms_lde_batch_scatter": (ci, [vp, ci, vp, sz, ui, ui, ui, u64, vp, sz, vp, sz, vp, sz]),
    "ms_ipc_export": (ci, [vp, vp, vp]),
    "ms_ipc_open": (ci, [vp, vp, C.POINTER(vp)]),
    "ms_ipc_close": (ci, [vp, vp]),
    "ms_scan_affine": (ci, [vp, ci, vp, ci, vp, vp, ci, sz, vp, ci, vp]),
    "ms_fri_fold": (ci, [vp, ci, vp, ui, ui, u64, vp, vp]),
    "ms_eval_constraints": (ci, [vp, vp, ui, vp, ui, vp, sz, ui, vp, sz, ui, ci, ui, u64, ci, ci, vp]),
    "ms_eval_constraints_ptrs": (ci, [vp, vp, ui, vp, ui, vp, vp, ui, ci, ui, u64, ci, ci, vp]),
    "ms_eval_jit_check": (ci, [vp, ui, vp, ui, ci, C.c_char_p, sz]),
    "ms_poly_eval": (ci, [vp, ci, vp, sz, ui, sz, vp, ui, vp]),
    "ms_fill_random": (ci, [vp, vp, sz, u64]),
}


def header_symbols():
    """every function name declared in include/ministark_b200.h"""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ms_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if b"sm_100a" not in lib.ms_version():      # only the CUDA build is ever used: there is no CPU path in the product
            raise RuntimeError(f"{LIB_PATH} is not the sm_100a build of libministark_b200 ({lib.ms_version()!r})")
        _lib = lib
    return _lib
