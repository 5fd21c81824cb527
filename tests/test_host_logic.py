"""Host side of the boundary on CPU: the drop-in modules keep the reference's constructor signatures, attribute
surface, state-dict keys and initialisation; CPU tensors are refused (no fallback); the DP bucketing is sane."""
import inspect

import pytest
import torch

from conftest import golden
from graph_augmented_vision_transformers_b200 import dp, modules, ops
from oracle import vit_oracle

REF_VIT_SIG = ["img_size", "patch_size", "in_chans", "num_classes", "embed_dim", "depth", "num_heads", "mlp_ratio",
               "qkv_bias", "drop_rate", "attn_drop_rate", "drop_path_rate"]          # vit.py:125-127


def test_constructor_signatures_match_reference():
    p = inspect.signature(modules.VisionTransformer.__init__).parameters
    assert list(p)[1:13] == REF_VIT_SIG
    assert [p[n].default for n in REF_VIT_SIG] == [224, 16, 3, 14, 768, 12, 12, 4.0, True, 0.0, 0.0, 0.0]
    for extra in ("graph_mode", "graph_k", "graph_every"):
        assert p[extra].kind is inspect.Parameter.KEYWORD_ONLY
    a = inspect.signature(modules.Attention.__init__).parameters                    # vit.py:42
    assert list(a)[1:] == ["dim", "num_heads", "qkv_bias", "attn_drop", "proj_drop"]
    assert (a["num_heads"].default, a["qkv_bias"].default) == (8, False)
    b = inspect.signature(modules.Block.__init__).parameters                        # vit.py:100-101
    assert list(b)[1:8] == ["dim", "num_heads", "mlp_ratio", "qkv_bias", "drop", "attn_drop", "drop_path"]


def test_state_dict_keys_and_seeded_init_match_reference():
    """graph_mode=None: same keys, same shapes and - under the reference's seed - the same weights."""
    g = golden("vit_b16_seed42")
    torch.manual_seed(42)
    m = modules.VisionTransformer()
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"])
    wsum = float(sum(p.double().abs().sum() for p in m.parameters()))
    assert abs(wsum - float(g["weight_abs_sum"])) < 1e-6 * float(g["weight_abs_sum"])
    torch.manual_seed(42)
    o = vit_oracle.VisionTransformer()
    assert list(m.state_dict()) == list(o.state_dict())


def test_graph_keys_are_additive():
    cfg = dict(img_size=32, patch_size=8, embed_dim=64, depth=4, num_heads=4)
    plain = set(modules.VisionTransformer(**cfg).state_dict())
    graph = set(modules.VisionTransformer(**cfg, graph_mode="knn", graph_k=4, graph_every=2).state_dict())
    extra = sorted(graph - plain)
    assert plain < graph
    assert extra == sorted(f"blocks.{i}.{n}" for i in (0, 2) for n in
                           ("norm_g.weight", "norm_g.bias", "graph.proj.weight", "graph.proj.bias"))
    # a reference checkpoint therefore loads with strict=False and leaves only graph keys missing
    msg = modules.VisionTransformer(**cfg, graph_mode="knn", graph_k=4, graph_every=2).load_state_dict(
        modules.VisionTransformer(**cfg).state_dict(), strict=False)
    assert sorted(msg.missing_keys) == extra and not msg.unexpected_keys


def test_attribute_surface_used_by_gradcam():
    blk = modules.Block(64, 4, qkv_bias=True)
    assert isinstance(blk.attn.qkv, torch.nn.Linear) and isinstance(blk.attn.proj, torch.nn.Linear)
    assert isinstance(blk.attn.attn_drop, torch.nn.Dropout) and isinstance(blk.norm1, torch.nn.LayerNorm)
    assert blk.attn.num_heads == 4 and blk.attn.scale == 16 ** -0.5
    with pytest.raises(AssertionError):
        modules.Attention(65, num_heads=4)                                            # vit.py:44


def test_load_mae_weights_skips_head(tmp_path):
    cfg = dict(img_size=32, patch_size=8, embed_dim=64, depth=1, num_heads=4)
    src = modules.VisionTransformer(**cfg)
    with torch.no_grad():
        for p in src.parameters():
            p.add_(1.0)
    path = tmp_path / "mae.pth"
    torch.save({"model": src.state_dict()}, path)
    dst = modules.VisionTransformer(**cfg, graph_mode="knn", graph_k=4)
    head_before = dst.head.weight.clone()
    dst.load_mae_weights(str(path))
    assert torch.equal(dst.blocks[0].attn.qkv.weight, src.blocks[0].attn.qkv.weight)
    assert torch.equal(dst.head.weight, head_before)


def test_cpu_tensors_are_refused_not_emulated():
    m = modules.VisionTransformer(img_size=32, patch_size=8, embed_dim=64, depth=1, num_heads=4, graph_mode="knn")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.attention_core(torch.randn(1, 4, 3 * 64), 1, 0.125)
    with pytest.raises(AssertionError):
        m.patch_embed(torch.randn(1, 3, 31, 32))                                      # vit.py:27-28


def test_grad_sync_buckets_reverse_order_and_cover_all_params():
    m = vit_oracle.VisionTransformer(img_size=32, patch_size=8, embed_dim=64, depth=2, num_heads=4)
    gs = dp.GradSync(m, bucket_mb=0.05)
    flat = [p for b in gs.buckets for p in b]
    params = [p for p in m.parameters()]
    assert [id(p) for p in flat] == [id(p) for p in reversed(params)]
    assert len(gs.buckets) > 2
    assert dp.shard_batch(2048, 3, 8) == (768, 1024)
    with pytest.raises(ValueError):
        dp.shard_batch(10, 0, 4)


def test_loss_matches_reference_fixture():
    from graph_augmented_vision_transformers_b200.losses import DynamicWeightedLoss
    g = golden("vit_small")
    crit = DynamicWeightedLoss(14)
    assert sorted(crit.state_dict()) == ["lambda_asl", "lambda_focal", "lambda_wbce", "pos_weight"]
    loss, parts = crit(torch.from_numpy(g["logits"]), torch.from_numpy(g["tgt"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    for n in ("wbce", "focal", "asl"):
        assert abs(float(parts[n]) - float(g[n])) < 1e-6, n


def test_ops_refuse_cpu_tensors_for_the_new_rows():
    """f2 / f4 host logic: no CPU fallback behind the token prologue, loud errors for unsupported patch sizes."""
    import torch
    from graph_augmented_vision_transformers_b200 import ops
    img, w = torch.randn(1, 3, 32, 32), torch.randn(64, 3, 16, 16)
    z = torch.zeros(1, 5, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.patch_embed_tokens(img, w, None, z[:, :1], z, 0.0, False)
    assert ops.patch_embed_supported(img, w)
    assert not ops.patch_embed_supported(img, torch.randn(64, 3, 12, 12))          # 12 is not a multiple of 8
    assert not ops.patch_embed_supported(torch.randn(1, 3, 40, 40), w)             # 40 is not a multiple of 16


def test_shadow_cache_ignores_cpu_parameters_and_follows_epochs():
    import torch
    from graph_augmented_vision_transformers_b200 import ops
    p = torch.nn.Parameter(torch.randn(4, 4))
    ops.refresh_shadows([p, None], torch.bfloat16)                                  # CPU parameters are skipped, not cast
    assert id(p) not in ops._SHADOWS
    s = ops._shadow(p, torch.bfloat16)                                              # one-off cast, no autograd edge
    assert s.dtype == torch.bfloat16 and not s.requires_grad and torch.equal(s, p.detach().bfloat16())
    assert ops._shadow(p, torch.float32).data_ptr() == p.data_ptr()                 # same dtype: the parameter itself
    e0 = ops._SHADOW_EPOCH["n"]
    ops.invalidate_shadows()
    assert ops._SHADOW_EPOCH["n"] == e0 + 1


def test_captured_step_requires_a_capturable_optimizer():
    import torch
    from graph_augmented_vision_transformers_b200.step import CapturedTrainStep
    m = torch.nn.Linear(4, 4)
    with pytest.raises(ValueError, match="capturable"):
        CapturedTrainStep(m, torch.nn.MSELoss(), torch.optim.AdamW(m.parameters()))
    with pytest.raises(ValueError, match="int64"):
        from graph_augmented_vision_transformers_b200 import ops
        ops.set_rng_offset_tensor(torch.zeros(1))


def test_colsum_workspace_covers_the_launchers_chunking():
    """ops._colsum_ws must be an upper bound of the row chunks the C launchers use (<= 12 per SM and column block)."""
    from graph_augmented_vision_transformers_b200 import ops
    import torch

    def chunks(rows, D, per_sm, sms=148):                                           # colsum_grid() of edges.cu
        cb = (D + 255) // 256
        n = min((per_sm * sms + cb - 1) // cb, 1024)
        if n * 32 > rows:
            n = (rows + 31) // 32
        rpc = (rows + n - 1) // n
        return (rows + rpc - 1) // rpc

    for rows, D in [(256 * 197, 3072), (256 * 197, 768), (256, 197 * 768), (2, 64), (50432, 2304)]:
        cb = (D + 255) // 256
        bound = min(1024, (12 * 148 + cb - 1) // cb + 1, max(1, (rows + 31) // 32) + 1)
        for per_sm in (6, 12):
            assert chunks(rows, D, per_sm) <= bound, (rows, D, per_sm)


def test_product_path_never_touches_the_oracle_or_the_reference_tree():
    """The oracle is test infrastructure: nothing under the package, src/ or the C sources may import it or read
    /root/reference; bench.py may (cpu_baseline / --impl reference legs only) and must not import it at module level."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for base in ("graph_augmented_vision_transformers_b200", "src", "include"):
        for dp, _, files in os.walk(os.path.join(root, base)):
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h")):
                    continue
                text = open(os.path.join(dp, f), errors="ignore").read()
                code = "\n".join(ln for ln in text.splitlines() if not ln.lstrip().startswith(("#", "//", "*", "/*")))
                if re.search(r"^\s*(from|import)\s+oracle\b", code, flags=re.M) or re.search(r"open\([^)]*/root/reference", code):
                    offenders.append(os.path.join(dp, f))
    assert not offenders, offenders
    bench = open(os.path.join(root, "bench.py")).read()
    top_level = [ln for ln in bench.splitlines() if re.match(r"(from|import)\s+oracle\b", ln)]
    assert not top_level, "bench.py must import the oracle only inside its cpu_baseline / reference legs"
