"""a1-a4 - multi-head self-attention, /root/reference/src/models/vit.py:39-72, through gvit_attn_fwd / gvit_attn_bwd."""
import pytest
import torch

from conftest import TOL_BF16, TOL_F32, golden, rel_err
from gpu_util import DEV
from graph_augmented_vision_transformers_b200 import _lib, modules, ops
from oracle import vit_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,dim", [("attn_small", 128), ("attn_dh64", 192)])
def test_fp32_module_matches_reference_fixture(name, dim):
    g = golden(name)
    m = modules.Attention(dim, num_heads=int(g["heads"]), qkv_bias=True).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.copy_(torch.from_numpy(g["param." + n]))
    m.to(DEV)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    out = m(x)
    out.backward(torch.from_numpy(g["cot"]).to(DEV))
    assert rel_err(out, g["out"]) < TOL_F32 and rel_err(x.grad, g["dx"]) < TOL_F32
    for n, p in m.named_parameters():
        assert rel_err(p.grad, g["grad." + n]) < TOL_F32, n


def _core_case(B, N, H, dh, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    qkv = torch.randn(B, N, 3 * H * dh, generator=g).to(dtype).float()
    cot = torch.randn(B, N, H * dh, generator=g).to(dtype).float()
    qd = qkv.to(DEV, dtype).requires_grad_(True)
    out = ops.attention_core(qd, H, dh ** -0.5)
    out.backward(cot.to(DEV, dtype))
    qc = qkv.clone().requires_grad_(True)
    want = vit_oracle.attention_core(qc, H, dh ** -0.5)
    want.backward(cot)
    return out, qd.grad, want, qc.grad


@pytest.mark.parametrize("B,N,H,dh", [(2, 197, 12, 64), (1, 577, 16, 64), (2, 37, 4, 32), (1, 1, 2, 64), (3, 128, 2, 64),
                                      (2, 129, 3, 64)])
def test_fp32_core_vs_oracle(B, N, H, dh):
    out, dqkv, want, dwant = _core_case(B, N, H, dh, torch.float32)
    assert rel_err(out, want) < TOL_F32 and rel_err(dqkv, dwant) < TOL_F32


@pytest.mark.parametrize("B,N,H", [(2, 197, 12), (1, 577, 16), (1, 1, 2), (3, 128, 2), (2, 129, 3), (2, 16, 1), (1, 256, 4),
                                   (1, 257, 2), (31, 197, 12), (40, 65, 8), (27, 150, 6),   # these three: > 148 (image, head)
                                                                                            # items -> persistent CTAs loop
                                   (2, 300, 3), (1, 640, 2), (1, 1025, 1), (3, 577, 16)])   # N > 256: two-pass tcgen05 backward
def test_bf16_core_vs_oracle(B, N, H):
    assert _lib.describe_path("attn_fwd", _lib.GVIT_BF16, N, 64) == "attn_fwd:tcgen05+tma"
    assert _lib.describe_path("attn_bwd", _lib.GVIT_BF16, N, 64) == "attn_bwd:tcgen05+tma"
    out, dqkv, want, dwant = _core_case(B, N, H, 64, torch.bfloat16, seed=N)
    assert out.dtype == torch.bfloat16
    assert rel_err(out, want) < TOL_BF16 and rel_err(dqkv, dwant) < TOL_BF16


def test_bf16_under_autocast_matches_module_semantics():
    m = modules.Attention(768, num_heads=12, qkv_bias=True).to(DEV)
    o = vit_oracle.Attention(768, num_heads=12, qkv_bias=True)
    o.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
    x = torch.randn(2, 197, 768, generator=torch.Generator().manual_seed(1))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        got = m(x.to(DEV))
    assert got.dtype == torch.bfloat16 and rel_err(got, o(x)) < TOL_BF16
    with torch.autocast("cuda", dtype=torch.float16):             # the reference trainer's autocast dtype
        got16 = m(x.to(DEV))
    assert rel_err(got16, o(x)) < TOL_BF16


def test_softmax_rows_sum_to_one_at_full_size():
    """config 2 size: with v == 1 every output is exactly the softmax row sum."""
    B, N, H = 256, 197, 12
    qkv = torch.randn(B, N, 3, H, 64, device=DEV, dtype=torch.bfloat16, generator=torch.Generator(device=DEV).manual_seed(0))
    qkv[:, :, 2] = 1.0
    out = ops.attention_core(qkv.view(B, N, -1), H, 0.125)
    assert (out.float() - 1).abs().max() < 1e-2
    # linearity in v
    qkv2 = qkv.clone()
    qkv2[:, :, 2] = 2.0
    assert (ops.attention_core(qkv2.view(B, N, -1), H, 0.125).float() - 2).abs().max() < 2e-2


def test_attn_drop_runs_the_reference_op_sequence_and_says_so():
    """attn_drop > 0 (vit.py:66; 0 in every shipped configuration): the fused kernels keep no (B,H,N,N) tensor, so training
    with it runs the reference's op sequence on the GPU - with a RuntimeWarning; eval mode stays on the fused kernel."""
    torch.manual_seed(0)
    m = modules.Attention(128, num_heads=2, qkv_bias=True, attn_drop=0.25).to(DEV)
    x = torch.randn(3, 50, 128, device=DEV)
    modules.Attention._warned_attn_drop = False
    m.train()
    with pytest.warns(RuntimeWarning, match="attn_drop"):
        torch.manual_seed(7)
        y1 = m(x)
    torch.manual_seed(7)
    y2 = m(x)
    assert torch.equal(y1, y2)                                  # same seed, same mask
    # the same mask through the plain reference formula (vit.py:59-71)
    torch.manual_seed(7)
    B, N, C = x.shape
    qkv = m.qkv(x).reshape(B, N, 3, 2, C // 2).permute(2, 0, 3, 1, 4)
    attn = torch.nn.functional.dropout(((qkv[0] @ qkv[1].transpose(-2, -1)) * m.scale).softmax(dim=-1), 0.25, True)
    want = m.proj((attn @ qkv[2]).transpose(1, 2).reshape(B, N, C))
    assert rel_err(y1, want) < TOL_F32
    xg = x.clone().requires_grad_(True)
    m(xg).square().sum().backward()
    assert torch.isfinite(xg.grad).all() and all(torch.isfinite(p.grad).all() for p in m.parameters())
    m.eval()
    m2 = modules.Attention(128, num_heads=2, qkv_bias=True, attn_drop=0.0).to(DEV).eval()
    m2.load_state_dict(m.state_dict())
    assert torch.equal(m(x), m2(x))                             # no dropout in eval: the fused kernel either way


def test_unsupported_head_dim_is_refused_loudly():
    with pytest.raises(_lib.GvitError, match="GVIT_ERR_UNSUPPORTED"):
        ops.attention_core(torch.randn(1, 4, 3 * 48, device=DEV), 1, 1.0)      # head dim 48


@pytest.mark.parametrize("N,boost_from,boost", [(197, 100, 60.0), (197, 33, 25.0), (256, 200, 200.0), (150, 64, 8.0),
                                                (577, 300, 60.0), (513, 420, 200.0), (300, 290, 25.0)])   # N > 256: the online-softmax
                                                # rescale of the key-block kernel (running max jumps in a LATER block)
def test_bf16_large_logits_late_keys(N, boost_from, boost):
    """Softmax range stress: keys from `boost_from` on are scaled so that the row maxima sit far (tens to hundreds of nats)
    above the scores of the first keys, in some heads and not in others; forward, saved log-sum-exp (through the
    backward) against the fp32 oracle."""
    B, H = 2, 3
    g = torch.Generator().manual_seed(N + int(boost))
    qkv = torch.randn(B, N, 3, H, 64, generator=g)
    qkv[:, boost_from:, 1] *= boost                              # later keys: logits of up to +-(4 * boost) nats
    qkv[0, :, 0, 0] *= 0.01                                      # one head with tiny queries: its rows never re-base
    qkv = qkv.bfloat16().float().reshape(B, N, 3 * H * 64)
    out = ops.attention_core(qkv.to(DEV).bfloat16(), H, 0.125)
    q, k, v = qkv.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    want = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, dim=-1) @ v).transpose(1, 2).reshape(B, N, H * 64)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, want) < TOL_BF16
    # the saved log-sum-exp must be the true one (the backward recomputes P from it): check through the gradient
    x = qkv.to(DEV).bfloat16().requires_grad_(True)
    o = ops.attention_core(x, H, 0.125)
    cot = torch.randn(B, N, H * 64, generator=g).bfloat16()
    o.backward(cot.to(DEV))
    xr = qkv.clone().requires_grad_(True)
    qr, kr, vr = xr.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    (torch.softmax((qr @ kr.transpose(-1, -2)) * 0.125, dim=-1) @ vr).transpose(1, 2).reshape(B, N, H * 64).backward(cot.float())
    assert rel_err(x.grad, xr.grad) < 2 * TOL_BF16              # near-one-hot rows: bf16 P / dS rounding on huge logits
