"""The C-ABI boundary without a GPU: libgvit.so loads on a CPU-only machine, exports every function that
include/gvit.h declares, the ctypes table mirrors the header, and argument validation fails loudly."""
import ctypes
import os
import re

import pytest

from graph_augmented_vision_transformers_b200 import _lib

HEADER = open(_lib.HEADER_PATH).read()
DECLS = re.findall(r"GVIT_API\s+[\w\s\*]+?\b(gvit_\w+)\s*\(([^;]*?)\)\s*;", HEADER, flags=re.S)


def test_header_declares_the_expected_surface():
    names = [n for n, _ in DECLS]
    for must in ("gvit_knn_fwd", "gvit_knn_bwd", "gvit_graph_reverse", "gvit_agg_fwd", "gvit_agg_gather_fwd",
                 "gvit_agg_bwd", "gvit_graph_bwd", "gvit_attn_fwd", "gvit_attn_bwd", "gvit_layernorm_fwd", "gvit_layernorm_bwd",
                 "gvit_colsum", "gvit_dropout_residual_fwd", "gvit_dropout_bwd", "gvit_gelu_dropout_fwd", "gvit_gelu_dropout_bwd",
                 "gvit_patchify", "gvit_embed_assemble", "gvit_version", "gvit_last_error_string"):
        assert must in names
    assert len(names) == len(set(names))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build libgvit.so first: make (or __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name, _ in DECLS:
        assert hasattr(lib, name), f"{name} is declared in gvit.h but not exported"


def test_ctypes_table_mirrors_header():
    assert set(_lib.SIGNATURES) == {n for n, _ in DECLS}
    for name, args in DECLS:
        args = args.strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        assert len(_lib.SIGNATURES[name]) == n, name


def test_version_and_constants():
    lib = _lib.load()
    assert lib.gvit_version() == _lib.ABI_VERSION == int(re.search(r"#define GVIT_ABI_VERSION (\d+)", HEADER).group(1))
    assert _lib.GVIT_MAX_K == int(re.search(r"GVIT_MAX_K = (\d+)", HEADER).group(1))
    assert _lib.GVIT_LN_PARTIALS == int(re.search(r"GVIT_LN_PARTIALS = (\d+)", HEADER).group(1))


def test_validation_errors_are_loud_and_need_no_gpu():
    # argument checks run before any CUDA call, so they are observable on a CPU-only machine
    with pytest.raises(_lib.GvitError, match="GVIT_ERR_DTYPE"):
        _lib.call("gvit_attn_fwd", 16, 1, 4, 1, 64, 0.125, 7, 16, 16, None)
    with pytest.raises(_lib.GvitError, match="GVIT_ERR_SHAPE"):
        _lib.call("gvit_knn_fwd", 16, 64, 64, 1, 4, 64, 9, _lib.GVIT_F32, 16, 16, 16, None)     # k > Np
    with pytest.raises(_lib.GvitError, match="GVIT_ERR_ALIGN"):
        _lib.call("gvit_layernorm_fwd", 8, 16, 16, 1, 64, 1e-5, _lib.GVIT_F32, _lib.GVIT_F32, 16, 16, 16, None)
    with pytest.raises(_lib.GvitError, match="GVIT_ERR_DTYPE"):            # bf16 stream with an fp32 branch is not a pairing
        _lib.call("gvit_layernorm_fwd", 16, 16, 16, 1, 64, 1e-5, _lib.GVIT_BF16, _lib.GVIT_F32, 16, 16, 16, None)
    with pytest.raises(_lib.GvitError, match="GVIT_ERR_UNSUPPORTED"):
        _lib.call("gvit_agg_fwd", 16, 1, 16, 64, 4, _lib.GVIT_F32, 16, 16, 16, None, None, _lib.GVIT_F32, 16, None, None, 0, None)
    with pytest.raises(_lib.GvitError, match="GVIT_ERR_SHAPE"):            # patch size must be a multiple of 8
        _lib.call("gvit_patchify", 16, 1, 3, 224, 224, 14, _lib.GVIT_F32, _lib.GVIT_BF16, 16, None)
    with pytest.raises(_lib.GvitError, match="GVIT_ERR_DTYPE"):            # a bf16 image cannot produce fp32 patches
        _lib.call("gvit_patchify", 16, 1, 3, 224, 224, 16, _lib.GVIT_BF16, _lib.GVIT_F32, 16, None)
    with pytest.raises(_lib.GvitError, match="GVIT_ERR_SHAPE"):            # dropout without a keep-mask buffer
        _lib.call("gvit_embed_assemble", 16, 16, 16, 16, 1, 197, 768, 0.1, 1, 0, None, _lib.GVIT_BF16, _lib.GVIT_F32, _lib.GVIT_BF16, 16, None, None)
    with pytest.raises(_lib.GvitError, match="GVIT_ERR_UNSUPPORTED"):      # the fused graph backward is bf16-only
        _lib.call("gvit_graph_bwd", 16, 1024, 64, 1, 16, 64, 4, _lib.GVIT_F32, 16, 16, 16, 16, 16, 1024, 16, 16, None)


def test_describe_path_routes_bf16_to_tcgen05():
    assert _lib.describe_path("knn", _lib.GVIT_BF16, 196, 768) == "knn:tcgen05+tma"
    assert _lib.describe_path("knn", _lib.GVIT_F32, 196, 768) == "knn:fp32-fma"
    assert _lib.describe_path("agg", _lib.GVIT_BF16, 196, 768) == "agg:tcgen05+tma"
    assert _lib.describe_path("attn_fwd", _lib.GVIT_BF16, 197, 64) == "attn_fwd:tcgen05+tma"
    assert _lib.describe_path("graph_bwd", _lib.GVIT_BF16, 196, 768) == "graph_bwd:tcgen05+tma"
    assert _lib.describe_path("graph_bwd", _lib.GVIT_BF16, 576, 1024) == "graph_bwd:reverse-csr+gather"
