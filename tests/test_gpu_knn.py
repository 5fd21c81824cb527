"""a7 - graph construction through gvit_knn_fwd.

fp32 (the exact-FMA kernel): neighbour indices AND similarities BIT-EXACT on EVERY row against the strict-order fp32
oracle (oracle/knn_strict.c, GRAPH_SPEC_VERSION 2) - north_star "kNN neighbour indices bit-exact in fp32 (ties broken by
lowest index)".  bf16 (the tcgen05 kernel): fp32 accumulation in the tensor core's own order, so indices are exact on
every row whose decision margin exceeds that round-off; the all-row match fraction against the strict oracle is printed
(run with -s) and asserted against the measured floor."""
import numpy as np
import pytest
import torch

from conftest import golden
from gpu_util import DEV, check_adjacency, tokens
from graph_augmented_vision_transformers_b200 import _lib, ops
from oracle import knn_strict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["graph_knn_small", "graph_knn_196", "graph_ties"])
def test_fp32_golden_indices_bit_exact(name):
    g = golden(name)
    assert int(g["spec_version"]) == 2
    h = torch.from_numpy(g["h"])
    idx, vals, _ = ops.knn_graph(h.to(DEV), int(g["k"]))
    assert np.array_equal(idx.cpu().numpy(), g["idx"])                 # every row, no margin filter
    assert np.array_equal(vals.cpu().numpy(), g["vals"])               # same operations in the same order: same bits


@pytest.mark.parametrize("tag,B,Np,D,k", [("b196", 2, 196, 768, 4), ("b196", 2, 196, 768, 8), ("b196", 2, 196, 768, 16),
                                          ("l576", 1, 576, 1024, 8)])
def test_fp32_benchmark_shapes_bit_exact_vs_committed_strict_adjacency(tag, B, Np, D, k):
    """BASELINE shapes (ViT-B/16: 196 tokens x 768, k in {4,8,16}; ViT-L/16 @ 384: 576 x 1024): the committed strict
    adjacency (tests/golden/graph_knn_strict.npz, made by oracle/make_golden.py) and a live strict-oracle run."""
    g = golden("graph_knn_strict")
    h = torch.randn(B, Np + 1, D, generator=torch.Generator().manual_seed(100 + Np))
    assert abs(float(h.double().abs().sum()) - float(g[tag + "_abs_sum"])) < 1e-9 * float(g[tag + "_abs_sum"])   # RNG drift guard
    idx, vals, rnorm = ops.knn_graph(h.to(DEV), k)
    assert np.array_equal(idx.cpu().numpy(), g[f"{tag}_k{k}_idx"].astype(np.int32))
    assert np.array_equal(vals.cpu().numpy(), g[f"{tag}_k{k}_vals"])
    li, lv, lr = knn_strict.knn_strict(h[:, 1:].numpy(), k)
    assert np.array_equal(idx.cpu().numpy(), li) and np.array_equal(vals.cpu().numpy(), lv)
    assert np.array_equal(rnorm.cpu().numpy(), lr)


@pytest.mark.parametrize("B,Np,D,k", [(3, 50, 72, 5), (1, 8, 8, 8), (2, 129, 192, 32), (4, 65, 256, 1), (2, 300, 128, 16)])
def test_fp32_ragged_shapes_bit_exact_vs_live_strict_oracle(B, Np, D, k):
    hc, hd = tokens(B, Np, D, seed=Np + k)
    hc[0, 3] = hc[0, 1]                                                # exact duplicates: the tie rule is exercised too
    hd = hc.to(DEV)
    idx, vals, rnorm = ops.knn_graph(hd, k)
    li, lv, lr = knn_strict.knn_strict(hc[:, 1:].numpy(), k)
    assert np.array_equal(idx.cpu().numpy(), li) and np.array_equal(vals.cpu().numpy(), lv)
    assert np.array_equal(rnorm.cpu().numpy(), lr)
    if k >= 2:
        assert list(idx[0, 0, :2].cpu()) == [0, 2] and list(idx[0, 2, :2].cpu()) == [0, 2]


def test_bf16_storage_fp32_arithmetic_kernel_is_bit_exact_too():
    """bf16-stored tokens outside every tcgen05 range (D % 64 != 0) run the same exact-FMA kernel on the bf16 values."""
    hc, hd = tokens(1, 300, 72, seed=3, dtype=torch.bfloat16)
    assert _lib.describe_path("knn", _lib.GVIT_BF16, 300, 72) == "knn:fp32-fma"
    idx, vals, _ = ops.knn_graph(hd, 8)
    li, lv, _ = knn_strict.knn_strict(hc[:, 1:].numpy(), 8)
    assert np.array_equal(idx.cpu().numpy(), li) and np.array_equal(vals.cpu().numpy(), lv)


def test_bf16_more_than_256_tokens_runs_on_tensor_cores():
    """256 < Np <= 1024 (the 576 patch tokens of a 384x384 image): Gram matrix by gvit_bgemm + gvit_knn_select; exact ties
    still resolve to the lowest index (rows 7 and 400 are duplicates)."""
    assert _lib.describe_path("knn", _lib.GVIT_BF16, 576, 1024) == "knn:tcgen05 gram + row select"
    hc, _ = tokens(2, 576, 256, seed=11, dtype=torch.bfloat16)
    hc[:, 401] = hc[:, 8]                                   # patch rows 7 and 400 identical
    idx, vals, rnorm = ops.knn_graph(hc.to(DEV, torch.bfloat16), 8)
    idx = idx.cpu().numpy()
    assert list(idx[0, 7, :2]) == [7, 400] and list(idx[0, 400, :2]) == [7, 400]
    assert (idx[..., 0][:, [i for i in range(576) if i != 400]] == np.array([i for i in range(576) if i != 400])).all()
    assert (np.diff(vals.cpu().numpy(), axis=-1) <= 0).all()


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


# all-row match fraction of the tcgen05 kernel against the strict fp32 oracle on the same bf16 values: MEASURED 1.00000 on
# every shape below (profiles/r2b_knn_match.txt); the floor only leaves room for a near-tie inside the accumulation-order
# noise on another seed.  (was 0.97 in round 1, never measured)
# measured floor of the all-row match fraction of the tcgen05 kernel against the strict fp32 oracle on the same bf16
# values (profiles/r2*_knn_match.txt); rows that differ are near-ties inside the accumulation-order noise
BF16_ALL_ROW_MATCH_FLOOR = 0.995


@pytest.mark.parametrize("B,Np,D,k", [(3, 196, 768, 8), (2, 196, 768, 4), (2, 196, 768, 16), (1, 16, 64, 1),
                                      (2, 256, 1024, 32), (2, 129, 192, 8), (1, 64, 128, 8), (2, 576, 1024, 8)])
def test_bf16_indices_vs_float64_oracle(B, Np, D, k):
    hc, hd = tokens(B, Np, D, seed=Np + k, dtype=torch.bfloat16)
    idx, vals, rnorm = ops.knn_graph(hd, k)
    check_adjacency(hc, idx, vals, k, noise=1e-5)
    li, lv, _ = knn_strict.knn_strict(hc[:, 1:].numpy(), k)
    same = float((idx.cpu().numpy() == li).all(-1).mean())
    print(f"\nknn bf16 ({B},{Np},{D},k={k}) path={_lib.describe_path('knn', _lib.GVIT_BF16, Np, D)}: rows identical to the "
          f"strict fp32 oracle {same:.5f}, max |dval| {np.abs(vals.cpu().numpy() - lv).max():.2e}")
    assert same >= (BF16_ALL_ROW_MATCH_FLOOR if k <= 16 else 0.9)
    want = 1.0 / hc[:, 1:].double().norm(dim=-1)
    assert float((rnorm.cpu().double() - want).abs().max() / want.max()) < 1e-5
    assert (idx[..., 0].cpu() == torch.arange(Np, dtype=torch.int32)).all()      # self loop first (S_ii = 1)


@pytest.mark.parametrize("B,Np,D,k", [(2, 196, 768, 8), (1, 50, 72, 5), (2, 576, 1024, 8), (1, 8, 8, 8)])
def test_fp32_indices_vs_float64_oracle(B, Np, D, k):
    """the strict fp32 result is also the mathematically right one wherever fp32 can tell (float64 cross-check).
    The noise band follows the specification's sequential fp32 chain: its round-off grows with the chain length D
    (measured max |S_fp32 - S_f64| = 2.03e-6 at D = 1024), so the band is 3e-9 * D, floored at 2e-6."""
    hc, hd = tokens(B, Np, D, seed=Np + k)
    idx, vals, _ = ops.knn_graph(hd, k)
    check_adjacency(hc, idx, vals, k, noise=max(2e-6, 3e-9 * D))


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
