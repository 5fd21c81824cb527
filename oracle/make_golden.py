"""Generate tests/golden/*.npz.  Run in the build container only:

    python -m oracle.make_golden

The attention / block / ViT / loss fixtures come from the *imported reference*
(/root/reference/src/models/vit.py, /root/reference/src/training/losses.py), not
from the restatement in oracle/vit_oracle.py - they are what pins the oracle.
The graph fixtures come from oracle/graph_oracle.py (the reference has no graph
code; they freeze GRAPH_SPEC_VERSION and are labelled "parity unpinned").
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def _np(t):
    return t.detach().cpu().numpy()


def _ref_modules():
    # the repo has its own `src` shim package; make sure the reference's wins here
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        from src.models import vit as ref_vit
        from src.training import losses as ref_losses
    finally:
        sys.path.remove(REF)
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    return ref_vit, ref_losses


def _randomise(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in mod.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.2 if p.ndim > 1 else 0.1))


def module_case(mod, x, cot, prefix=""):
    x = x.clone().requires_grad_(True)
    out = mod(x)
    out.backward(cot)
    d = {prefix + "x": _np(x), prefix + "cot": _np(cot), prefix + "out": _np(out), prefix + "dx": _np(x.grad)}
    for n, p in mod.named_parameters():
        d[prefix + "param." + n] = _np(p)
        d[prefix + "grad." + n] = _np(p.grad)
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if "--graph-only" not in sys.argv:
        reference_fixtures()
    graph_fixtures()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def reference_fixtures():
    ref_vit, ref_losses = _ref_modules()
    g = torch.Generator().manual_seed(0)

    # 1. Attention (vit.py:39-72): small dims and one head_dim=64 case with a ragged N
    for name, (dim, heads, N) in {"attn_small": (128, 4, 37), "attn_dh64": (192, 3, 197)}.items():
        m = ref_vit.Attention(dim, num_heads=heads, qkv_bias=True).eval()
        _randomise(m, 1)
        x = torch.randn(2, N, dim, generator=g)
        cot = torch.randn(2, N, dim, generator=g)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), heads=heads, **module_case(m, x, cot))

    # 2. Block (vit.py:97-119)
    m = ref_vit.Block(128, num_heads=4, mlp_ratio=2.0, qkv_bias=True).eval()
    _randomise(m, 2)
    with torch.no_grad():
        for n, p in m.named_parameters():          # keep LayerNorm near (1, 0)
            if "norm" in n and n.endswith("weight"):
                p.mul_(0.5).add_(1.0)
    x = torch.randn(2, 37, 128, generator=g)
    cot = torch.randn(2, 37, 128, generator=g)
    np.savez_compressed(os.path.join(OUT, "block_small.npz"), heads=4, mlp_ratio=2.0, **module_case(m, x, cot))

    # 3. Small VisionTransformer + DynamicWeightedLoss (vit.py:122-224, losses.py:7-68)
    torch.manual_seed(42)                                              # scripts/train.py:137
    cfg = dict(img_size=32, patch_size=8, in_chans=3, num_classes=14, embed_dim=64, depth=2, num_heads=4,
               mlp_ratio=2.0, qkv_bias=True)
    m = ref_vit.VisionTransformer(**cfg).eval()
    crit = ref_losses.DynamicWeightedLoss(14)
    img = torch.randn(3, 3, 32, 32, generator=g)
    tgt = (torch.rand(3, 14, generator=g) > 0.7).float()
    logits = m(img)
    loss, parts = crit(logits, tgt)
    loss.backward()
    d = {"img": _np(img), "tgt": _np(tgt), "logits": _np(logits), "loss": _np(loss),
         "wbce": _np(parts["wbce"]), "focal": _np(parts["focal"]), "asl": _np(parts["asl"])}
    for n, p in m.named_parameters():
        d["param." + n] = _np(p)
        d["grad." + n] = _np(p.grad)
    np.savez_compressed(os.path.join(OUT, "vit_small.npz"), **{k: v for k, v in d.items()},
                        cfg=np.array(repr(cfg)))

    # 4. Full ViT-B/16 under the reference's seed: logits only; weights/inputs are regenerated from the
    #    seed by the test and guarded by checksums so RNG drift is detected instead of mis-reported.
    torch.manual_seed(42)
    m = ref_vit.VisionTransformer().eval()
    gi = torch.Generator().manual_seed(1234)
    img = torch.randn(2, 3, 224, 224, generator=gi)
    with torch.no_grad():
        logits = m(img)
        feats = m.forward_features(img)
    wsum = float(sum(p.double().abs().sum() for p in m.parameters()))
    np.savez_compressed(os.path.join(OUT, "vit_b16_seed42.npz"), logits=_np(logits), feats_head=_np(feats[:, :16]),
                        weight_abs_sum=wsum, img_abs_sum=float(img.double().abs().sum()),
                        n_params=sum(p.numel() for p in m.parameters()))


def graph_fixtures():
    # 5. Graph layer - from oracle/graph_oracle.py, PARITY UNPINNED.  idx / vals are the STRICT fp32
    #    evaluation (oracle/knn_strict.c, spec version 2); out / gradients come from autograd over that adjacency.
    from oracle import GRAPH_SPEC_VERSION
    from oracle.graph_oracle import graph_layer_forward, knn_select_strict
    gg = torch.Generator().manual_seed(7)
    for name, (Np, D, k, mode) in {"graph_knn_small": (20, 32, 4, "knn"), "graph_dense_small": (20, 32, 0, "dense"),
                                   "graph_knn_196": (196, 64, 8, "knn")}.items():
        h = torch.randn(2, Np + 1, D, generator=gg).requires_grad_(True)
        W = (torch.randn(D, D, generator=gg) * 0.2).requires_grad_(True)
        b = (torch.randn(D, generator=gg) * 0.1).requires_grad_(True)
        cot = torch.randn(2, Np + 1, D, generator=gg)
        out, aux = graph_layer_forward(h, W, b, k, mode, return_aux=True)
        out.backward(cot)
        d = dict(spec_version=GRAPH_SPEC_VERSION, h=_np(h), W=_np(W), b=_np(b), cot=_np(cot), out=_np(out),
                 dh=_np(h.grad), dW=_np(W.grad), db=_np(b.grad), k=k, mode=np.array(mode))
        if mode == "knn":
            si, sv = knn_select_strict(h[:, 1:], k)
            assert torch.equal(si, aux["idx"])
            d.update(idx=_np(si).astype(np.int32), vals=_np(sv))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)

    # 6. exact-tie case: duplicated rows must resolve to the lowest index
    h = torch.randn(1, 33, 16, generator=gg)
    h[0, 1 + 9] = h[0, 1 + 5]
    h[0, 1 + 17] = h[0, 1 + 5]
    h[0, 1 + 1] = h[0, 1 + 0]
    si, sv = knn_select_strict(h[:, 1:], 4)
    np.savez_compressed(os.path.join(OUT, "graph_ties.npz"), spec_version=GRAPH_SPEC_VERSION, h=_np(h),
                        idx=_np(si).astype(np.int32), vals=_np(sv), k=4)
    # 7. strict adjacency at the benchmark shapes (inputs are regenerated from the seed by the tests; only idx is stored):
    #    (2,196,768) for k in {4,8,16} and (1,576,1024) k = 8
    d = {}
    for tag, (B, Np, D, ks) in {"b196": (2, 196, 768, (4, 8, 16)), "l576": (1, 576, 1024, (8,))}.items():
        hh = torch.randn(B, Np + 1, D, generator=torch.Generator().manual_seed(100 + Np))
        d[tag + "_abs_sum"] = float(hh.double().abs().sum())
        for k in ks:
            si, sv = knn_select_strict(hh[:, 1:], k)
            d[f"{tag}_k{k}_idx"] = _np(si).astype(np.int16)
            d[f"{tag}_k{k}_vals"] = _np(sv)
    np.savez_compressed(os.path.join(OUT, "graph_knn_strict.npz"), spec_version=GRAPH_SPEC_VERSION, **d)


if __name__ == "__main__":
    main()
