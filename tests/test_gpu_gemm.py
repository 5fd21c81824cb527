"""a1 / f1 - the nn.Linear GEMMs (vit.py:59 qkv, :70 proj, :90/:93 Mlp, :28 patch projection) through gvit_linear_gemm
(2-SM tcgen05 kernel, csrc/gemm2_tc.cu): forward, input gradient, weight gradient against an fp32 product of the same
bf16 operands (north_star: 2e-2 relative in bf16), and the properties that hold at full size."""
import pytest
import torch
import torch.nn.functional as F

from conftest import TOL_BF16, rel_err
from gpu_util import DEV
from graph_augmented_vision_transformers_b200 import _lib, ops

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).bfloat16().to(DEV)


# rows: a multiple of 256, a ragged tail, fewer rows than one tile, the bench shape
@pytest.mark.parametrize("M", [512, 1000, 77, 256 * 197])
@pytest.mark.parametrize("K,N", [(768, 2304), (768, 768), (3072, 768), (1024, 4096), (64, 256)])
def test_forward_matches_fp32_product(M, K, N):
    if M * max(K, N) > 60e6 and (K, N) not in ((768, 2304), (3072, 768)):
        pytest.skip("full-size rows on the two bench shapes only")
    x, w, b = _rand((M, K), 1), _rand((N, K), 2, K ** -0.5), _rand((N,), 3)
    y = ops._mm_nt(x, w, b)
    assert y.dtype == torch.bfloat16 and y.shape == (M, N)
    want = x.float() @ w.float().t() + b.float()
    assert rel_err(y, want) < TOL_BF16 / 4
    assert torch.equal(ops._mm_nt(x, w, b), y), "the forward GEMM is not run-to-run deterministic"
    y3 = ops._mm_nt(x.view(1, M, K), w, None)                              # (B, N, K) input, no bias
    assert y3.shape == (1, M, N) and rel_err(y3[0], x.float() @ w.float().t()) < TOL_BF16 / 4


@pytest.mark.parametrize("M", [512, 1000, 256 * 197])
@pytest.mark.parametrize("K,N", [(768, 2304), (768, 768), (3072, 768), (768, 3072)])
def test_input_and_weight_gradient(M, K, N):
    if M * max(K, N) > 60e6 and (K, N) != (768, 2304):
        pytest.skip("full-size rows on the qkv shape only")
    x, w, dy = _rand((M, K), 4), _rand((N, K), 5, K ** -0.5), _rand((M, N), 6)
    dx = ops._mm_nn(dy, w)
    assert dx.dtype == torch.bfloat16 and rel_err(dx, dy.float() @ w.float()) < TOL_BF16 / 4
    dw = ops._wgrad(dy, x, torch.float32)
    assert dw.dtype == torch.float32 and dw.shape == (N, K)
    assert rel_err(dw, dy.float().t() @ x.float()) < 1e-4                 # fp32 accumulation of exact bf16 products
    assert torch.equal(ops._wgrad(dy, x, torch.float32), dw), "split-K pieces must be added in a fixed order"


def test_weight_gradient_split_is_used_and_linear():
    """At the bench shape the reduction (50432 rows) is cut into pieces; the result must still be the plain sum: linear in
    dy, and equal to the sum of the gradients of two row halves."""
    M, K, N = 256 * 197, 768, 768
    assert _lib.load().gvit_linear_gemm_ws_bytes(N, K, M) > 0
    x, dy = _rand((M, K), 7), _rand((M, N), 8)
    dw = ops._wgrad(dy, x, torch.float32)
    h = (M // 2) // 256 * 256
    parts = ops._wgrad(dy[:h], x[:h], torch.float32) + ops._wgrad(dy[h:], x[h:], torch.float32)
    assert rel_err(dw, parts) < 1e-5
    assert rel_err(ops._wgrad(dy * 2, x, torch.float32), 2 * dw) < 1e-6   # powers of two scale exactly in bf16


def test_linear_autograd_uses_the_kernel():
    """ops.linear under autocast: y, dx, dW, db against torch's F.linear on the same bf16 operands."""
    M, K, N = 4 * 197, 768, 2304
    x = _rand((4, 197, K), 9).float().requires_grad_(True)
    w = (torch.randn(N, K, generator=torch.Generator().manual_seed(10)) * K ** -0.5).to(DEV).requires_grad_(True)
    b = torch.zeros(N, device=DEV, requires_grad=True)
    cot = _rand((4, 197, N), 11)
    ops.reset_launch_count()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = ops.linear(x, w, b)
    y.backward(cot)
    assert ops.launch_count() >= 4                                          # three GEMMs + the bias column sum
    xr, wr, br = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    yr = F.linear(xr.bfloat16().float(), wr.bfloat16().float(), br)
    yr.backward(cot.float())
    assert rel_err(y, yr) < TOL_BF16 and rel_err(x.grad, xr.grad) < TOL_BF16
    assert w.grad.dtype == torch.float32 and rel_err(w.grad, wr.grad) < TOL_BF16 and rel_err(b.grad, br.grad) < TOL_BF16


def test_unsupported_shapes_are_refused_not_miscomputed():
    x, w = _rand((64, 768), 12), _rand((200, 768), 13)
    out = torch.empty(64, 200, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(_lib.GvitError):
        _lib.call("gvit_linear_gemm", x.data_ptr(), 0, 768, w.data_ptr(), 0, 768, 64, 200, 768, None, _lib.GVIT_BF16, out.data_ptr(), 200,
                  None, 0, torch.cuda.current_stream().cuda_stream)
    assert rel_err(ops._mm_nt(x, w), x.float() @ w.float().t()) < TOL_BF16  # the host side routes N % 256 != 0 to the library GEMM
