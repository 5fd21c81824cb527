"""a5 - LayerNorm (nn.LayerNorm at vit.py:103,108,154) and proj_drop + residual (vit.py:71,117)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import TOL_BF16, TOL_F32, rel_err
from gpu_util import DEV
from graph_augmented_vision_transformers_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,D", [(2 * 197, 768), (577, 1024), (5, 64), (1, 8), (256 * 197, 768)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm_forward_backward(rows, D, dtype):
    g = torch.Generator().manual_seed(rows)
    x = (torch.randn(rows, D, generator=g) * 2 + 0.5).to(dtype).float()
    w = (1 + 0.2 * torch.randn(D, generator=g)).to(dtype).float()
    b = (0.1 * torch.randn(D, generator=g)).to(dtype).float()
    cot = torch.randn(rows, D, generator=g).to(dtype).float()
    xd, wd, bd = (t.to(DEV, dtype).requires_grad_(True) for t in (x, w, b))
    y = ops.layer_norm(xd, wd, bd, 1e-5)
    y.backward(cot.to(DEV, dtype))
    xc, wc, bc = (t.clone().requires_grad_(True) for t in (x, w, b))
    F.layer_norm(xc, (D,), wc, bc, 1e-5).backward(cot)
    tol = TOL_F32 if dtype == torch.float32 else TOL_BF16
    assert rel_err(y, F.layer_norm(x, (D,), w, b, 1e-5)) < tol
    assert rel_err(xd.grad, xc.grad) < tol and rel_err(wd.grad, wc.grad) < tol and rel_err(bd.grad, bc.grad) < tol


def test_dropout_add():
    x = torch.randn(64, 197, 768, device=DEV)
    r = torch.randn_like(x)
    assert torch.equal(ops.dropout_add(x, r, 0.1, training=False), x + r)
    assert ops.dropout_add(x, None, 0.0, training=True) is x
    torch.manual_seed(0)
    xr = x.clone().requires_grad_(True)
    y = ops.dropout_add(xr, r, 0.1, training=True)
    kept = (y != r)
    assert abs(float(kept.float().mean()) - 0.9) < 2e-3                      # Bernoulli(0.9) keep mask
    assert rel_err(y[kept], (x / 0.9 + r)[kept]) < 1e-6
    y.backward(torch.ones_like(y))
    assert torch.equal(xr.grad != 0, kept) and rel_err(xr.grad[kept], torch.full_like(xr.grad[kept], 1 / 0.9)) < 1e-6
    torch.manual_seed(0)
    assert torch.equal(ops.dropout_add(x, r, 0.1, training=True), y)          # reproducible under manual_seed
    assert not torch.equal(ops.dropout_add(x, r, 0.1, training=True), y)      # and fresh on the next call
    m = (ops.dropout_add(torch.ones(8, 1 << 20, device=DEV), None, 0.5, True) != 0).float()
    assert abs(float(m.mean()) - 0.5) < 2e-3 and abs(float((m[:, 1:] * m[:, :-1]).mean()) - 0.25) < 2e-3
    yb = ops.dropout_add(x.bfloat16(), r, 0.0, training=True)                 # bf16 branch onto an fp32 stream
    assert yb.dtype == torch.float32 and rel_err(yb, x.bfloat16().float() + r) < 1e-6
