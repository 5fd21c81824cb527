"""a5/a6 - Block and VisionTransformer (vit.py:97-224) assembled from the libgvit-backed modules, against the
reference fixtures (fp32) and the oracle graph-ViT (fp32 and bf16 autocast)."""
import ast

import pytest
import torch

from conftest import TOL_BF16, TOL_F32, golden, rel_err
from gpu_util import DEV
from graph_augmented_vision_transformers_b200 import modules, ops
from oracle import vit_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _exact_fp32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _load(mod, g, prefix="param."):
    with torch.no_grad():
        for n, p in mod.named_parameters():
            p.copy_(torch.from_numpy(g[prefix + n]))
    return mod.to(DEV)


def test_block_matches_reference_fixture():
    g = golden("block_small")
    m = _load(modules.Block(128, num_heads=int(g["heads"]), mlp_ratio=float(g["mlp_ratio"]), qkv_bias=True).eval(), g)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    out = m(x)
    out.backward(torch.from_numpy(g["cot"]).to(DEV))
    assert rel_err(out, g["out"]) < TOL_F32 and rel_err(x.grad, g["dx"]) < TOL_F32
    for n, p in m.named_parameters():
        assert rel_err(p.grad, g["grad." + n]) < TOL_F32, n


def test_vit_and_loss_match_reference_fixture():
    g = golden("vit_small")
    m = _load(modules.VisionTransformer(**ast.literal_eval(str(g["cfg"]))).eval(), g)
    logits = m(torch.from_numpy(g["img"]).to(DEV))
    assert rel_err(logits, g["logits"]) < TOL_F32
    loss = vit_oracle.multilabel_loss(logits, torch.from_numpy(g["tgt"]).to(DEV), torch.ones(3, device=DEV),
                                      torch.ones(14, device=DEV))
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    for n, p in m.named_parameters():
        assert rel_err(p.grad, g["grad." + n]) < 2 * TOL_F32, n


def test_vit_b16_seed42_logits_match_reference():
    g = golden("vit_b16_seed42")
    torch.manual_seed(42)
    m = modules.VisionTransformer().eval().to(DEV)
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        assert rel_err(m(img.to(DEV)), g["logits"]) < TOL_F32
        assert rel_err(m.forward_features(img.to(DEV))[:, :16], g["feats_head"]) < TOL_F32


CFG = dict(img_size=64, patch_size=8, embed_dim=128, depth=2, num_heads=2, mlp_ratio=2.0, graph_mode="knn", graph_k=4)


def _pair(cfg=CFG, seed=0):
    torch.manual_seed(seed)
    o = vit_oracle.VisionTransformer(**cfg).eval()
    m = modules.VisionTransformer(**cfg).eval()
    m.load_state_dict(o.state_dict())
    return o, m.to(DEV)


def test_graph_vit_fp32_forward_backward_vs_oracle():
    o, m = _pair()
    g = torch.Generator().manual_seed(1)
    img, tgt = torch.randn(3, 3, 64, 64, generator=g), (torch.rand(3, 14, generator=g) > 0.7).float()
    lo = vit_oracle.multilabel_loss(o(img), tgt, torch.ones(3), torch.ones(14))
    logits = m(img.to(DEV))
    lm = vit_oracle.multilabel_loss(logits, tgt.to(DEV), torch.ones(3, device=DEV), torch.ones(14, device=DEV))
    assert rel_err(logits, o(img)) < TOL_F32 and abs(float(lo) - float(lm)) < 1e-5
    lo.backward()
    lm.backward()
    for (n, a), b in zip(o.named_parameters(), m.parameters()):
        assert rel_err(b.grad, a.grad) < 3 * TOL_F32, n


def _pin_adjacency(o, m, img_dev, autocast):
    """Run the device model once with graph recording on and pin the oracle's per-layer adjacency to the one the device
    built from ITS tokens of that layer.  Deep in a network a near-tie neighbour can flip on 1e-7 of input noise, and a
    flipped neighbour is not a small perturbation of the output; what must hold layer by layer is (a) identical inputs
    -> identical indices (checked here for fp32 against the strict oracle, bit-exact on every row) and (b) identical
    indices -> activations / gradients within tolerance (checked by the callers on logits and every gradient)."""
    from oracle import knn_strict
    for blk in m.blocks:
        if hasattr(blk, "graph"):
            blk.graph.record_graph = True
    with torch.no_grad():
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                m(img_dev)
        else:
            m(img_dev)
    n_checked = 0
    for bo, bm in zip(o.blocks, m.blocks):
        if hasattr(bm, "graph"):
            bo.graph.idx_override = bm.graph.last_idx.cpu().long()
            if not autocast:
                li, lv, _ = knn_strict.knn_strict(bm.graph.last_tokens[:, 1:].float().cpu().numpy(), bm.graph.k)
                assert (bm.graph.last_idx.cpu().numpy() == li).all(), "fp32 kNN indices differ from the strict oracle"
                assert (bm.graph.last_vals.cpu().numpy() == lv).all()
                n_checked += 1
            bm.graph.record_graph = False
            bm.graph.last_tokens = bm.graph.last_idx = bm.graph.last_vals = None
    return n_checked


def _worst_grad_errors(o, m):
    errs = {n: rel_err(b.grad, a.grad) for (n, a), b in zip(o.named_parameters(), m.parameters())}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    return errs, worst


VITB_GRAPH = dict(img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12, graph_mode="knn", graph_k=8)


def test_vit_b16_graph_depth12_fp32_logits_and_every_gradient():
    """BASELINE configs[1] model (ViT-B/16 + kNN graph block in every layer, depth 12), batch 4, fp32: logits and EVERY
    gradient - graph.proj.*, norm_g.* included - within north_star's 1e-4, kNN indices bit-exact in all 12 layers."""
    o, m = _pair(VITB_GRAPH, seed=42)
    g = torch.Generator().manual_seed(1234)
    img, tgt = torch.randn(4, 3, 224, 224, generator=g), (torch.rand(4, 14, generator=g) > 0.9).float()
    assert _pin_adjacency(o, m, img.to(DEV), autocast=False) == 12
    want = o(img)
    vit_oracle.multilabel_loss(want, tgt, torch.ones(3), torch.ones(14)).backward()
    logits = m(img.to(DEV))
    vit_oracle.multilabel_loss(logits, tgt.to(DEV), torch.ones(3, device=DEV), torch.ones(14, device=DEV)).backward()
    errs, worst = _worst_grad_errors(o, m)
    print(f"\nfp32 depth-12 graph-ViT-B: logits rel err {rel_err(logits, want):.2e}; worst gradients {worst}")
    assert rel_err(logits, want) < TOL_F32
    assert any("graph.proj" in n for n in errs) and any("norm_g" in n for n in errs)
    for n, e in errs.items():
        assert e < TOL_F32, (n, e)


@pytest.mark.parametrize("fp32_residual", [True, False])
def test_vit_b16_graph_depth12_bf16_autocast_logits_and_every_gradient(fp32_residual):
    """Same model under bf16 autocast (the bench's precision), per-layer adjacency pinned to the device's: logits and
    every gradient within north_star's 2e-2 of the fp32 oracle.  fp32_residual=True is torch.autocast's own semantics
    (the default, and what bench.py's `value` measures); False is the opt-in bf16 residual stream."""
    o, m = _pair(VITB_GRAPH, seed=42)
    m.fp32_residual = fp32_residual
    g = torch.Generator().manual_seed(1234)
    img, tgt = torch.randn(4, 3, 224, 224, generator=g), (torch.rand(4, 14, generator=g) > 0.9).float()
    _pin_adjacency(o, m, img.to(DEV), autocast=True)
    want = o(img)
    vit_oracle.multilabel_loss(want, tgt, torch.ones(3), torch.ones(14)).backward()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = m(img.to(DEV))
    vit_oracle.multilabel_loss(logits.float(), tgt.to(DEV), torch.ones(3, device=DEV), torch.ones(14, device=DEV)).backward()
    errs, worst = _worst_grad_errors(o, m)
    print(f"\nbf16 depth-12 graph-ViT-B (fp32_residual={fp32_residual}): logits rel err {rel_err(logits, want):.2e}; "
          f"worst gradients {worst}")
    # fp32_residual=True (default, what `value` measures): north_star's 2e-2 on logits and on EVERY gradient.
    # fp32_residual=False is the explicit opt-in that narrows the residual stream below what torch.autocast gives the
    # reference: its extra rounding drift is the named exception (3e-2, measured worst 2.34e-2 on blocks.4.norm1.bias).
    tol = TOL_BF16 if fp32_residual else 1.5 * TOL_BF16
    assert rel_err(logits, want) < tol
    for n, e in errs.items():
        assert e < tol, (n, e)


def test_parameter_shadows_follow_data_writes():
    """`.data` writes do not move `_version`; an autocast forward after one must still see the new weights.
    Each forward enters its own autocast region, as a training loop does (torch.autocast's OWN weight cache - which the
    classifier head, a plain nn.Linear, goes through - lives for one region; that is torch behaviour, reference included)."""
    _, m = _pair(seed=8)
    img = torch.randn(2, 3, 64, 64, device=DEV)

    def fwd(mod):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return mod(img)

    a = fwd(m)
    for p in m.parameters():
        p.data.mul_(1.5)                           # e.g. EMA / clamping / dist.broadcast(p.data): `_version` unchanged
    b = fwd(m)
    m2 = modules.VisionTransformer(**CFG).eval().to(DEV)
    m2.load_state_dict(m.state_dict())
    c = fwd(m2)
    assert not torch.equal(a, b) and torch.equal(b, c)


def test_graph_every_and_dense_variants_fp32():
    for extra in (dict(graph_every=2), dict(graph_mode="dense")):
        o, m = _pair({**CFG, **extra}, seed=3)
        img = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(2))
        with torch.no_grad():
            assert rel_err(m(img.to(DEV)), o(img)) < TOL_F32


def test_graph_vit_bf16_autocast_trains():
    """bf16 autocast fwd+bwd (the bench's precision): close to the fp32 oracle, finite grads for every parameter."""
    o, m = _pair(seed=5)
    g = torch.Generator().manual_seed(1)
    img, tgt = torch.randn(4, 3, 64, 64, generator=g), (torch.rand(4, 14, generator=g) > 0.7).float()
    _pin_adjacency(o, m, img.to(DEV), autocast=True)       # near-tie neighbour swaps cannot hide (or cause) errors
    m.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = m(img.to(DEV))
    assert rel_err(logits, o(img)) < TOL_BF16
    loss = vit_oracle.multilabel_loss(logits.float(), tgt.to(DEV), torch.ones(3, device=DEV), torch.ones(14, device=DEV))
    before = ops.launch_count()
    loss.backward()
    assert ops.launch_count() > before                     # the backward ran through libgvit kernels
    for n, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    lo = vit_oracle.multilabel_loss(o(img), tgt, torch.ones(3), torch.ones(14))
    lo.backward()
    errs, worst = _worst_grad_errors(o, m)                 # every parameter, graph.* and norm_g.* included
    print(f"\nbf16 small graph-ViT: worst gradients {worst}")
    assert worst[0][1] < TOL_BF16, worst


def test_pure_bf16_model_runs():
    _, m = _pair(seed=6)
    m.bfloat16()
    out = m(torch.randn(2, 3, 64, 64, device=DEV, dtype=torch.bfloat16))
    assert out.dtype == torch.bfloat16 and torch.isfinite(out.float()).all()


def test_fp32_residual_stream_option_and_hook_safe_folding():
    o, m = _pair(seed=7)
    img = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    assert m.fp32_residual is True                         # default = torch.autocast's own residual-stream semantics
    _pin_adjacency(o, m, img.to(DEV), autocast=True)
    with torch.autocast("cuda", dtype=torch.bfloat16), torch.no_grad():
        a = m(img.to(DEV))
    m.fp32_residual = False
    _pin_adjacency(o, m, img.to(DEV), autocast=True)
    with torch.autocast("cuda", dtype=torch.bfloat16), torch.no_grad():
        b = m(img.to(DEV))
    assert rel_err(b, o(img)) < TOL_BF16
    m.fp32_residual = True
    _pin_adjacency(o, m, img.to(DEV), autocast=True)
    assert rel_err(a, o(img)) < TOL_BF16
    for bo in o.blocks:
        bo.graph.idx_override = None
    # a forward hook on blocks.1.attn (what Grad-CAM installs) must observe the attention output, not x + attention
    seen = {}
    h = m.blocks[1].attn.register_forward_hook(lambda mod, inp, out: seen.setdefault("out", out.detach()))
    with torch.no_grad():
        hooked = m(img.to(DEV))
    h.remove()
    with torch.no_grad():
        plain = m(img.to(DEV))
    assert rel_err(hooked, plain) < 1e-5
    ref = {}
    ho = o.blocks[1].attn.register_forward_hook(lambda mod, inp, out: ref.setdefault("out", out.detach()))
    o(img)
    ho.remove()
    assert rel_err(seen["out"], ref["out"]) < TOL_F32


def _train_setup(drop, seed=0):
    from graph_augmented_vision_transformers_b200.losses import DynamicWeightedLoss
    torch.manual_seed(seed)
    cfg = dict(CFG, drop_rate=drop)
    m = modules.VisionTransformer(**cfg).to(DEV).train()
    crit = DynamicWeightedLoss(14).to(DEV)
    params = list(m.parameters()) + list(crit.parameters())
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.05, fused=True, capturable=True)
    return m, crit, opt, params


def test_captured_train_step_matches_the_eager_step():
    """f2 - trainer.py:96-120 as one CUDA graph: without dropout the replayed steps follow the eager trajectory."""
    from graph_augmented_vision_transformers_b200.step import CapturedTrainStep
    g = torch.Generator().manual_seed(2)
    batches = [(torch.randn(4, 3, 64, 64, generator=g).to(DEV), (torch.rand(4, 14, generator=g) > 0.7).float().to(DEV))
               for _ in range(4)]
    m0, c0, o0, p0 = _train_setup(0.0)
    eager = []
    # eager reference: 3 warm-up steps on batch 0 (what capture() does), then the 4 batches
    for img, tgt in [batches[0]] * 3 + batches:
        o0.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = m0(img)
        loss, _ = c0(logits, tgt)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(p0, 1.0, foreach=True)
        o0.step()
        eager.append(float(loss))
    m1, c1, o1, p1 = _train_setup(0.0)
    step = CapturedTrainStep(m1, c1, o1, max_norm=1.0, clip_params=p1, warmup=3)
    try:
        got = [float(step(img, tgt)) for img, tgt in batches]
        assert step.replays == 4
        for a, b in zip(got, eager[3:]):
            assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, eager[3:])
        for (n, a), b in zip(m1.named_parameters(), m0.parameters()):
            assert rel_err(a, b) < 2e-2, n
        # an eager evaluation after the replays must see the CURRENT weights (shadows are invalidated by every replay)
        m0.eval(), m1.eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            assert rel_err(m1(batches[0][0]), m0(batches[0][0])) < TOL_BF16
    finally:
        step.release()


def test_captured_step_draws_fresh_dropout_masks_on_every_replay():
    x = torch.ones(1 << 16, device=DEV)
    off = torch.zeros(1, dtype=torch.int64, device=DEV)
    ops.set_rng_offset_tensor(off)
    try:
        torch.manual_seed(0)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ops.dropout_add(x, None, 0.5, True)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            off.add_(1 << 40)
            y = ops.dropout_add(x, None, 0.5, True)
        masks = []
        for _ in range(3):
            graph.replay()
            masks.append((y != 0).clone())
        assert not torch.equal(masks[0], masks[1]) and not torch.equal(masks[1], masks[2])
        for mk in masks:
            assert abs(float(mk.float().mean()) - 0.5) < 1e-2
        assert abs(float((masks[0] & masks[1]).float().mean()) - 0.25) < 1e-2      # independent across replays
    finally:
        ops.set_rng_offset_tensor(None)


@pytest.mark.parametrize("k,every", [(4, 1), (16, 1), (8, 4)])
def test_config5_inference_sweep_point_vs_oracle(k, every):
    """BASELINE config 5 (inference: eval, no_grad, bf16, k in {4,8,16}, graph layers in every block vs every 4th) at
    ViT-B/16 width with 4 blocks: against the CPU oracle, and batch-size independence of the result."""
    cfg = dict(img_size=224, patch_size=16, embed_dim=768, depth=4, num_heads=12, graph_mode="knn", graph_k=k, graph_every=every)
    o, m = _pair(cfg, seed=k + every)
    img = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(k))
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        got = m(img.to(DEV))
        assert rel_err(got, o(img)) < 5e-2                 # whole-network bf16 drift (incl. near-tie neighbour swaps)
        big = torch.randn(64, 3, 224, 224, generator=torch.Generator(device=DEV).manual_seed(1), device=DEV)
        whole, parts = m(big), torch.cat([m(big[:32]), m(big[32:])])
    assert torch.equal(whole, parts)                       # images are independent: no batch-size dependent kernel path
    assert sum(hasattr(b, "graph") for b in m.blocks) == (4 if every == 1 else 1)


def test_config4_vit_l_384_dense_shapes_run():
    """BASELINE config 4 shapes (ViT-L/16 @ 384: 576 patch tokens, D = 1024, H = 16, dense adjacency), 2 blocks: the
    577-token attention runs on the tcgen05 forward, LayerNorm at D = 1024, finite gradients for every parameter."""
    cfg = dict(img_size=384, patch_size=16, embed_dim=1024, depth=2, num_heads=16, graph_mode="dense")
    o, m = _pair(cfg, seed=9)
    m.train()
    img = torch.randn(2, 3, 384, 384, generator=torch.Generator().manual_seed(4))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = m(img.to(DEV))
    assert rel_err(logits, o(img)) < TOL_BF16
    from graph_augmented_vision_transformers_b200 import _lib
    assert _lib.describe_path("agg_dense", _lib.GVIT_BF16, 576, 1024).startswith("agg_dense:tcgen05")
    assert _lib.describe_path("attn_bwd", _lib.GVIT_BF16, 577, 64) == "attn_bwd:tcgen05+tma"
    logits.float().square().mean().backward()
    o(img).square().mean().backward()
    errs, worst = _worst_grad_errors(o, m)
    print(f"\nconfig 4 (ViT-L/16 @ 384 dense, 2 blocks) bf16: worst gradients {worst}")
    for n, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    assert worst[0][1] < TOL_BF16, worst


def test_captured_forward_matches_eager_inference():
    """step.CapturedForward (the evaluate.py:104-115 call replayed from a CUDA graph, cached per batch shape) returns exactly
    what the eager bf16 forward returns, for two batch shapes, and refuses a model in training mode."""
    from graph_augmented_vision_transformers_b200.step import CapturedForward
    _, m = _pair(seed=12)
    fwd = CapturedForward(m)
    for B in (3, 8, 3):
        img = torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(B)).to(DEV)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            want = torch.sigmoid(m(img).float())
        assert torch.equal(fwd(img), want)
    assert len(fwd._graphs) == 2
    m.train()
    with pytest.raises(RuntimeError, match="eval"):
        fwd(img)


def test_repeated_batch128_inference_agg3_phase_guard():
    """Regression (r2y): at batch 128 the fused aggregation kernel launches a second wave of CTAs; an mt = 0 CTA whose
    warpgroup 1 was delayed by its global round trips waited on a barrier phase that warpgroup 0 had already completed twice
    and hung (~1 CTA in 10^5).  Thirty ViT-B/16 + k=4 graph forwards at that size, eager and replayed from a CUDA graph."""
    from graph_augmented_vision_transformers_b200.step import CapturedForward
    torch.manual_seed(42)
    m = modules.VisionTransformer(img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12, graph_mode="knn", graph_k=4).to(DEV).eval()
    img = torch.randn(128, 3, 224, 224, device=DEV)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ref = torch.sigmoid(m(img).float())
        for _ in range(14):
            out = torch.sigmoid(m(img).float())
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    fwd = CapturedForward(m)
    for _ in range(15):
        out = fwd(img)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
