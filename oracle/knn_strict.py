"""TEST INFRASTRUCTURE ONLY - Python binding of ``oracle/knn_strict.c`` (strict fp32 G1-G3, GRAPH_SPEC_VERSION 2).

**Parity unpinned** (the reference has no graph code): this is the builder-authored specification
of the fp32 accumulation order, written in C because only C's ``fmaf`` / ``sqrtf`` / ``/`` give single correctly
rounded binary32 operations (numpy would double-round through float64).  The shared object is built on first use
with gcc into ``oracle/_build/`` (git-ignored; it travels to the GPU box with the snapshot, and gcc is there too).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "knn_strict.c")
SO = os.path.join(_HERE, "_build", "libknn_strict.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off (no contraction of anything we did not write as fmaf) -> oracle/_build/libknn_strict.so."""
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        base = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", SO + ".tmp", "-lm"]
        try:                                     # hardware FMA when the host has it (fmaf is exact either way)
            subprocess.run(base[:3] + ["-march=native"] + base[3:], check=True, capture_output=True)
        except (subprocess.CalledProcessError, FileNotFoundError):
            subprocess.run(base, check=True)
        os.replace(SO + ".tmp", SO)
    return SO


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.knn_strict_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.knn_strict_f32.restype = C.c_int
        _lib = lib
    return _lib


def knn_strict(p, k: int):
    """p: (B, Np, D) array-like, taken as fp32.  Returns (idx int32 (B,Np,k), vals fp32 (B,Np,k), rnorm fp32 (B,Np))."""
    p = np.ascontiguousarray(np.asarray(p, dtype=np.float32))
    if p.ndim != 3:
        raise ValueError(f"p must be (B, Np, D); got {p.shape}")
    B, Np, D = p.shape
    idx = np.empty((B, Np, k), dtype=np.int32)
    vals = np.empty((B, Np, k), dtype=np.float32)
    rnorm = np.empty((B, Np), dtype=np.float32)
    rc = _load().knn_strict_f32(p.ctypes.data, B, Np, D, int(k), idx.ctypes.data, vals.ctypes.data, rnorm.ctypes.data)
    if rc != 0:
        raise ValueError(f"knn_strict_f32 rejected B={B} Np={Np} D={D} k={k}")
    return idx, vals, rnorm
