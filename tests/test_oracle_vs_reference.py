"""Live check of the oracle against the reference's own modules (only where /root/reference exists, i.e. in the
build container; on the GPU box the committed fixtures of test_oracle_golden.py carry the same evidence)."""
import os
import sys

import pytest
import torch

from conftest import rel_err
from oracle import vit_oracle

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "models")), reason="reference not mounted")


def _ref():
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "src" or k.startswith("src.")}
    sys.path.insert(0, REF)
    try:
        from src.models import vit as ref_vit
        from src.training import losses as ref_losses
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ref_vit, ref_losses


def test_vit_forward_backward_and_loss_match_reference():
    ref_vit, ref_losses = _ref()
    cfg = dict(img_size=48, patch_size=8, num_classes=14, embed_dim=96, depth=3, num_heads=3, mlp_ratio=2.0)
    torch.manual_seed(42)
    r = ref_vit.VisionTransformer(**cfg).eval()
    torch.manual_seed(42)
    o = vit_oracle.VisionTransformer(**cfg).eval()
    assert list(r.state_dict()) == list(o.state_dict())
    for (n, a), b in zip(r.state_dict().items(), o.state_dict().values()):
        assert torch.equal(a, b), n                       # same seed -> bit-identical initialisation
    img = torch.randn(4, 3, 48, 48)
    tgt = (torch.rand(4, 14) > 0.8).float()
    crit = ref_losses.DynamicWeightedLoss(14)
    lr, _ = crit(r(img), tgt)
    lo = vit_oracle.multilabel_loss(o(img), tgt, torch.ones(3), torch.ones(14))
    assert abs(float(lr) - float(lo)) < 1e-6
    lr.backward()
    lo.backward()
    for (n, a), b in zip(r.named_parameters(), o.parameters()):
        assert rel_err(b.grad, a.grad) < 1e-5, n


def test_attention_core_matches_reference_lines_59_69():
    ref_vit, _ = _ref()
    torch.manual_seed(0)
    m = ref_vit.Attention(192, num_heads=3, qkv_bias=True).eval()
    x = torch.randn(2, 50, 192)
    want = m(x)
    got = vit_oracle.attention_forward(x, m.qkv.weight, m.qkv.bias, m.proj.weight, m.proj.bias, 3)
    assert rel_err(got, want) < 1e-6
