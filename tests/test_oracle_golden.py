"""The CPU oracle against the committed fixtures.  Attention / Block / ViT / loss fixtures were produced by the
imported reference modules (oracle/make_golden.py), so these tests pin oracle/vit_oracle.py to the reference.
Graph fixtures freeze the SURVEY.md section 9 specification (parity unpinned: the reference has no graph code)."""
import ast

import numpy as np
import pytest
import torch

from conftest import golden, rel_err
from oracle import GRAPH_SPEC_VERSION, graph_oracle, vit_oracle


def _load_params(mod, g, prefix="param."):
    with torch.no_grad():
        for n, p in mod.named_parameters():
            p.copy_(torch.from_numpy(g[prefix + n]))


def _check_module(mod, g, tol=2e-6):
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    out = mod(x)
    out.backward(torch.from_numpy(g["cot"]))
    assert rel_err(out, g["out"]) < tol
    assert rel_err(x.grad, g["dx"]) < tol
    for n, p in mod.named_parameters():
        assert rel_err(p.grad, g["grad." + n]) < 5 * tol, n


@pytest.mark.parametrize("name,dim", [("attn_small", 128), ("attn_dh64", 192)])
def test_attention_matches_reference_fixture(name, dim):
    g = golden(name)
    m = vit_oracle.Attention(dim, num_heads=int(g["heads"]), qkv_bias=True).eval()
    _load_params(m, g)
    _check_module(m, g)


def test_block_matches_reference_fixture():
    g = golden("block_small")
    m = vit_oracle.Block(128, num_heads=int(g["heads"]), mlp_ratio=float(g["mlp_ratio"]), qkv_bias=True).eval()
    _load_params(m, g)
    _check_module(m, g)


def test_vit_and_loss_match_reference_fixture():
    g = golden("vit_small")
    cfg = ast.literal_eval(str(g["cfg"]))
    m = vit_oracle.VisionTransformer(**cfg).eval()
    _load_params(m, g)
    logits = m(torch.from_numpy(g["img"]))
    assert rel_err(logits, g["logits"]) < 2e-6
    loss = vit_oracle.multilabel_loss(logits, torch.from_numpy(g["tgt"]), torch.ones(3), torch.ones(14))
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    loss.backward()
    for n, p in m.named_parameters():
        assert rel_err(p.grad, g["grad." + n]) < 2e-5, n


def test_vit_b16_seed42_init_and_logits():
    """Same seed => same weights as the reference's initialisation (RNG order of vit.py:162-180) and same logits."""
    g = golden("vit_b16_seed42")
    torch.manual_seed(42)
    m = vit_oracle.VisionTransformer().eval()
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"]) == 85_809_422
    wsum = float(sum(p.double().abs().sum() for p in m.parameters()))
    assert abs(wsum - float(g["weight_abs_sum"])) < 1e-6 * float(g["weight_abs_sum"])
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1234))
    assert abs(float(img.double().abs().sum()) - float(g["img_abs_sum"])) < 1e-6 * float(g["img_abs_sum"])
    with torch.no_grad():
        logits = m(img)
    assert rel_err(logits, g["logits"]) < 1e-5


@pytest.mark.parametrize("name", ["graph_knn_small", "graph_dense_small", "graph_knn_196"])
def test_graph_spec_frozen(name):
    g = golden(name)
    assert int(g["spec_version"]) == GRAPH_SPEC_VERSION
    h = torch.from_numpy(g["h"]).requires_grad_(True)
    W = torch.from_numpy(g["W"]).requires_grad_(True)
    b = torch.from_numpy(g["b"]).requires_grad_(True)
    out, aux = graph_oracle.graph_layer_forward(h, W, b, int(g["k"]), str(g["mode"]), return_aux=True)
    out.backward(torch.from_numpy(g["cot"]))
    assert rel_err(out, g["out"]) < 2e-6
    assert rel_err(h.grad, g["dh"]) < 2e-5 and rel_err(W.grad, g["dW"]) < 2e-5 and rel_err(b.grad, g["db"]) < 2e-5
    assert float(out[:, 0].abs().max()) == 0.0                      # G0: CLS row untouched
    if str(g["mode"]) == "knn":
        assert np.array_equal(aux["idx"].numpy().astype(np.int32), g["idx"])       # strict fp32 order (spec version 2)
        assert np.abs(aux["vals"].detach().numpy() - g["vals"]).max() < 2e-6       # torch's S vs the strict chain: round-off
        idx64, _, margin = graph_oracle.knn_f64(g["h"][:, 1:], int(g["k"]))
        sure = margin > 1e-5                                        # rows whose answer fp32 noise cannot flip
        assert sure.mean() > 0.9
        assert np.array_equal(idx64[sure], g["idx"][sure])


def test_graph_ties_resolve_to_lowest_index():
    g = golden("graph_ties")
    h = torch.from_numpy(g["h"])
    _, aux = graph_oracle.graph_layer_forward(h, torch.eye(h.shape[-1]), None, int(g["k"]), "knn", return_aux=True)
    idx = aux["idx"].numpy()
    assert np.array_equal(idx.astype(np.int32), g["idx"])
    # rows 5, 9, 17 are identical: each of them lists the three in ascending order first
    for r in (5, 9, 17):
        assert list(idx[0, r, :3]) == [5, 9, 17]
    for r in (0, 1):
        assert list(idx[0, r, :2]) == [0, 1]
    idx64, _, _ = graph_oracle.knn_f64(g["h"][:, 1:], int(g["k"]))
    assert list(idx64[0, 9, :3]) == [5, 9, 17]


def test_strict_knn_known_answers():
    """oracle/knn_strict.c on cases whose answer is known in closed form."""
    from oracle import knn_strict
    # orthonormal rows: S = I exactly -> self first, then ties (all zeros) in ascending column order
    idx, vals, rn = knn_strict.knn_strict(np.eye(6, dtype=np.float32)[None] * 3.0, 3)
    for i in range(6):
        assert list(idx[0, i]) == [i] + [j for j in range(6) if j != i][:2]
    assert np.array_equal(vals[0, :, 0], np.ones(6, np.float32)) and np.all(vals[0, :, 1:] == 0)
    assert np.allclose(rn, 1 / 3.0)
    # collinear rows: cosine exactly +-1 regardless of scale; opposite direction sorts last
    p = np.array([[[1, 2, 2, 0], [2, 4, 4, 0], [-1, -2, -2, 0], [0, 0, 0, 5]]], dtype=np.float32)
    idx, vals, _ = knn_strict.knn_strict(p, 4)
    assert list(idx[0, 0]) == [0, 1, 3, 2] and list(idx[0, 1]) == [0, 1, 3, 2] and list(idx[0, 2]) == [2, 3, 0, 1]
    assert np.allclose(vals[0, 0], [1, 1, 0, -1], atol=1e-6)
    # zero row: norm clamps to 1e-12 (F.normalize), similarities 0, ties ascending
    p = np.zeros((1, 3, 8), np.float32)
    p[0, 1, 0] = 1.0
    idx, vals, rn = knn_strict.knn_strict(p, 3)
    assert list(idx[0, 0]) == [0, 1, 2] and np.all(vals[0, 0] == 0) and rn[0, 0] == np.float32(1e12)


def test_strict_knn_agrees_with_torch_and_float64_where_fp32_can_tell():
    """The strict order is one admissible fp32 evaluation of section 9: same indices as torch's GEMM-ordered fp32 and as
    float64 on every row whose decision margin exceeds fp32 round-off; exactly symmetric; committed fixture reproduced."""
    from oracle import knn_strict
    g = golden("graph_knn_strict")
    h = torch.randn(2, 197, 768, generator=torch.Generator().manual_seed(100 + 196))
    assert abs(float(h.double().abs().sum()) - float(g["b196_abs_sum"])) < 1e-9 * float(g["b196_abs_sum"])
    p = h[:, 1:]
    idx, vals, _ = knn_strict.knn_strict(p.numpy(), 8)
    assert np.array_equal(idx, g["b196_k8_idx"].astype(np.int32)) and np.array_equal(vals, g["b196_k8_vals"])
    ti, tv = graph_oracle.knn_select(graph_oracle.similarity(p), 8)
    idx64, _, margin = graph_oracle.knn_f64(p.numpy(), 8)
    sure = margin > 2e-6
    assert sure.mean() > 0.95
    assert np.array_equal(idx[sure], idx64[sure]) and np.array_equal(ti.numpy()[sure], idx64[sure])
    assert np.abs(vals - tv.numpy()).max() < 3e-6
    # symmetry of the strict S: i lists j with the same bits as j lists i
    look = {(0, i, int(j)): v for i in range(196) for j, v in zip(idx[0, i], vals[0, i])}
    pairs = [(k, v) for k, v in look.items() if (0, k[2], k[1]) in look]
    assert len(pairs) > 196 and all(look[(0, k[2], k[1])] == v for k, v in pairs)


def test_dense_is_the_k_equals_np_limit():
    torch.manual_seed(3)
    h, W, b = torch.randn(2, 13, 16), torch.randn(16, 16) * 0.2, torch.randn(16) * 0.1
    dense = graph_oracle.graph_layer_forward(h, W, b, 0, "dense")
    knn = graph_oracle.graph_layer_forward(h, W, b, 12, "knn")
    assert rel_err(knn, dense) < 1e-5
