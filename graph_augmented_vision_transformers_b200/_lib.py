"""ctypes binding of ``libgvit.so`` (the C ABI declared in ``include/gvit.h``).

The library is the product: there is no PyTorch/CPU fallback behind it.  If the
shared object is missing the import of any op raises ``GvitLibraryError`` telling
the user to build it (``make`` or ``python -c 'import __graft_entry__ as g; g.build()'``).
"""
from __future__ import annotations

import ctypes as C
import functools
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# GVIT_LIB=<path>: load another build of the library (same ABI version) - same-box A/B runs of bench.py / the tools
LIB_PATH = os.environ.get("GVIT_LIB") or os.path.join(_HERE, "lib", "libgvit.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "gvit.h")

GVIT_F32, GVIT_BF16 = 0, 1
GVIT_MAX_K = 32
GVIT_LN_PARTIALS = 296
GVIT_COLSUM_CHUNKS = 1024
ABI_VERSION = 18

STATUS_NAMES = {0: "GVIT_OK", 1: "GVIT_ERR_SHAPE", 2: "GVIT_ERR_ALIGN", 3: "GVIT_ERR_DTYPE", 4: "GVIT_ERR_CUDA",
                5: "GVIT_ERR_UNSUPPORTED"}


class GvitLibraryError(RuntimeError):
    pass


class GvitError(RuntimeError):
    """A libgvit entry point returned a non-zero gvit_status."""

    def __init__(self, fn, status, message):
        self.status = status
        super().__init__(f"{fn} failed with {STATUS_NAMES.get(status, status)}: {message}")


_vp, _i, _i64, _u64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float

# name -> argtypes; return type is int unless listed in _RESTYPE.  Order mirrors include/gvit.h.
SIGNATURES = {
    "gvit_version": [],
    "gvit_last_error_string": [],
    "gvit_describe_path": [C.c_char_p, _i, _i, _i, C.c_char_p, _i],
    "gvit_knn_fwd": [_vp, _i64, _i64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "gvit_graph_reverse": [_vp, _i, _i, _i, _vp, _vp, _vp],
    "gvit_knn_bwd": [_vp, _i64, _i64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "gvit_agg_gather_fwd": [_vp, _i64, _i64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "gvit_agg_fwd": [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i64, _vp],
    "gvit_agg_bwd": [_vp, _i64, _i64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "gvit_graph_bwd": [_vp, _i64, _i64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "gvit_bgemm": [_i, _i, _i, _i, _vp, _i64, _i64, _i, _vp, _i64, _i64, _i, _i, _vp, _i64, _i64, _i, _vp, _i64, _i64, _i, _i,
                   _vp, _i, _vp, _i64, _i64, _vp],
    "gvit_dense_rownorm": [_vp, _i64, _i64, _i, _i, _i, _vp, _vp],
    "gvit_knn_select": [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp],
    "gvit_dense_softmax_fwd": [_vp, _i, _vp, _i, _i, _i, _vp, _vp],
    "gvit_dense_softmax_bwd": [_vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp],
    "gvit_dense_combine_bwd": [_vp, _vp, _vp, _i64, _i64, _vp, _i, _i, _i, _vp, _vp],
    "gvit_attn_fwd": [_vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp],
    "gvit_attn_bwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp],
    "gvit_layernorm_fwd": [_vp, _vp, _vp, _i64, _i, _f, _i, _i, _vp, _vp, _vp, _vp],
    "gvit_layernorm_bwd": [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "gvit_colsum": [_vp, _i64, _i, _i, _i, _vp, _vp, _vp],
    "gvit_dropout_residual_fwd": [_vp, _vp, _i64, _f, _u64, _u64, _vp, _i, _i, _vp, _vp, _vp],
    "gvit_dropout_bwd": [_vp, _vp, _i64, _f, _i, _i, _vp, _i, _i, _vp, _vp, _vp],
    "gvit_gelu_dropout_fwd": [_vp, _i64, _f, _u64, _u64, _vp, _i, _vp, _vp, _vp],
    "gvit_gelu_dropout_bwd": [_vp, _vp, _vp, _i64, _f, _i, _vp, _i, _vp, _vp, _vp],
    "gvit_linear_gelu_dropout_fwd": [_vp, _vp, _vp, _i64, _i, _i, _f, _u64, _u64, _vp, _i, _i, _vp, _vp, _vp, _vp],
    "gvit_linear_dropout_residual_fwd": [_vp, _vp, _vp, _vp, _i64, _i, _i, _f, _u64, _u64, _vp, _i, _i, _vp, _vp, _vp],
    "gvit_linear_gelu_dropout_bwd_ws_rows": [_i64],
    "gvit_linear_gelu_dropout_bwd": [_vp, _vp, _vp, _vp, _i64, _i, _i, _f, _i, _i, _vp, _vp, _vp, _vp],
    "gvit_linear_gemm_ws_bytes": [_i64, _i, _i],
    "gvit_linear_gemm": [_vp, _i, _i64, _vp, _i, _i64, _i64, _i, _i, _vp, _i, _vp, _i64, _vp, _i64, _vp],
    "gvit_mt_chunk_elems": [],
    "gvit_mt_adamw_step": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _i64, _i64, _f, _f, _f, _vp, _vp, _vp, _vp],
    "gvit_patchify": [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "gvit_embed_assemble": [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _u64, _u64, _vp, _i, _i, _i, _vp, _vp, _vp],
}
_RESTYPE = {"gvit_last_error_string": C.c_char_p, "gvit_linear_gelu_dropout_bwd_ws_rows": C.c_int64,
             "gvit_linear_gemm_ws_bytes": C.c_int64}

_lib = None
_lock = threading.Lock()


def load():
    """Load libgvit.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise GvitLibraryError(
                f"{LIB_PATH} is missing: the CUDA library has not been built. Run `make` at the repository root "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no fallback path.")
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here == header/library out of sync
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, C.c_int)
        v = lib.gvit_version()
        if v != ABI_VERSION:
            raise GvitLibraryError(f"libgvit ABI version {v} != binding version {ABI_VERSION}; rebuild with `make`")
        _lib = lib
    return _lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.gvit_last_error_string()
        raise GvitError(name, rc, msg.decode() if msg else "")


@functools.lru_cache(maxsize=None)
def describe_path(op: str, dtype: int, n_tokens: int, dim: int) -> str:
    """Which kernels a request routes to (a pure function of its arguments: cached, the operators ask on every call)."""
    buf = C.create_string_buffer(64)
    call("gvit_describe_path", op.encode(), dtype, n_tokens, dim, buf, 64)
    return buf.value.decode()
