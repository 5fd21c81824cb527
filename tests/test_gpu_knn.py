"""a7 - graph construction through gvit_knn_fwd: neighbour indices bit-exact
(ties -> lowest index), similarities to fp32 round-off.  fp32 runs the exact-FMA kernel, bf16 the tcgen05 kernel."""
import numpy as np
import pytest
import torch

from conftest import golden
from gpu_util import DEV, check_adjacency, tokens
from graph_augmented_vision_transformers_b200 import _lib, ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["graph_knn_small", "graph_knn_196"])
def test_fp32_golden_indices_exact(name):
    g = golden(name)
    h = torch.from_numpy(g["h"])
    idx, vals, _ = ops.knn_graph(h.to(DEV), int(g["k"]))
    check_adjacency(h, idx, vals, int(g["k"]), noise=1e-6, min_sure=0.97)
    same = (idx.cpu().numpy() == g["idx"]).all(-1).mean()
    assert same > 0.97                                     # fixture came from torch CPU fp32 (its own round-off)
    assert np.abs(vals.cpu().numpy() - g["vals"]).max() < 2e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_exact_ties_resolve_to_lowest_index(dtype):
    g = golden("graph_ties")
    h = torch.from_numpy(g["h"]).to(dtype)
    if dtype == torch.bfloat16:                            # tcgen05 path needs D % 64 == 0: tile the features
        h = h.repeat(1, 1, 4)
    idx, vals, _ = ops.knn_graph(h.to(DEV), int(g["k"]))
    idx = idx.cpu().numpy()
    for r in (5, 9, 17):
        assert list(idx[0, r, :3]) == [5, 9, 17]
    for r in (0, 1):
        assert list(idx[0, r, :2]) == [0, 1]
    if dtype == torch.float32:
        assert np.array_equal(idx, g["idx"])


def test_duplicated_rows_everywhere_bf16():
    """every token duplicated once (rows i and i+98 identical): neighbour lists must come in ascending pairs."""
    hc, _ = tokens(2, 98, 128, seed=5, dtype=torch.bfloat16)
    h = torch.cat([hc, hc[:, 1:]], dim=1).to(torch.bfloat16)
    assert _lib.describe_path("knn", _lib.GVIT_BF16, 196, 128) == "knn:tcgen05+tma"
    idx, vals, _ = ops.knn_graph(h.to(DEV), 8)
    idx, vals = idx.cpu().numpy(), vals.cpu().numpy()
    assert (idx[..., 0::2] + 98 == idx[..., 1::2]).all()
    assert (vals[..., 0::2] == vals[..., 1::2]).all()


@pytest.mark.parametrize("B,Np,D,k", [(3, 196, 768, 8), (2, 196, 768, 4), (2, 196, 768, 16), (1, 16, 64, 1),
                                      (2, 256, 1024, 32), (2, 129, 192, 8), (1, 64, 128, 8), (2, 576, 1024, 8)])
def test_bf16_indices_vs_float64_oracle(B, Np, D, k):
    hc, hd = tokens(B, Np, D, seed=Np + k, dtype=torch.bfloat16)
    idx, vals, rnorm = ops.knn_graph(hd, k)
    check_adjacency(hc, idx, vals, k, noise=1e-5)
    want = 1.0 / hc[:, 1:].double().norm(dim=-1)
    assert float((rnorm.cpu().double() - want).abs().max() / want.max()) < 1e-5
    assert (idx[..., 0].cpu() == torch.arange(Np, dtype=torch.int32)).all()      # self loop first (S_ii = 1)


@pytest.mark.parametrize("B,Np,D,k", [(2, 196, 768, 8), (1, 50, 72, 5), (2, 576, 1024, 8), (1, 8, 8, 8)])
def test_fp32_indices_vs_float64_oracle(B, Np, D, k):
    hc, hd = tokens(B, Np, D, seed=Np + k)
    idx, vals, _ = ops.knn_graph(hd, k)
    check_adjacency(hc, idx, vals, k, noise=2e-6)


def test_full_size_properties_config2():
    """BASELINE config 2 size (batch 256, 196 tokens, D 768, k 8): size-independent properties."""
    g = torch.Generator(device=DEV).manual_seed(1)
    h = torch.randn(256, 197, 768, generator=g, device=DEV, dtype=torch.bfloat16)
    idx, vals, _ = ops.knn_graph(h, 8)
    ar = torch.arange(196, device=DEV, dtype=torch.int32)
    assert (idx[..., 0] == ar).all() and (vals[..., 0] - 1).abs().max() < 1e-5
    assert (vals[..., 1:] <= vals[..., :-1]).all() and (idx >= 0).all() and (idx < 196).all()
    assert (idx.sort(-1).values.diff(dim=-1) > 0).all()                          # no neighbour listed twice
    perm = torch.randperm(256, device=DEV)
    idx2, vals2, _ = ops.knn_graph(h[perm].contiguous(), 8)                      # images are independent
    assert torch.equal(idx2, idx[perm]) and torch.equal(vals2, vals[perm])
    # token permutation equivariance (checked through the similarity values, which are order-free)
    tp = torch.cat([torch.zeros(1, dtype=torch.long, device=DEV), 1 + torch.randperm(196, device=DEV)])
    idx3, vals3, _ = ops.knn_graph(h[:8, tp].contiguous(), 8)
    assert (vals3 - vals[:8, tp[1:] - 1]).abs().max() < 1e-5


def test_errors_are_loud():
    h = torch.randn(1, 17, 64, device=DEV)
    with pytest.raises(_lib.GvitError, match="GVIT_ERR_SHAPE"):
        ops.knn_graph(h, 17)
    with pytest.raises(TypeError):
        ops.knn_graph(h.half(), 4)
