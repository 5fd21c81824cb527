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
    y.backward(torch.ones_like(y))
    kept = xr.grad != 0                                                        # the keep mask, seen through the gradient
    assert abs(float(kept.float().mean()) - 0.9) < 2e-3                        # Bernoulli(0.9)
    assert rel_err(xr.grad[kept], torch.full_like(xr.grad[kept], 1 / 0.9)) < 1e-6
    assert rel_err(y[kept], (x / 0.9 + r)[kept]) < 1e-6 and torch.equal(y[~kept], r[~kept])
    torch.manual_seed(0)
    assert torch.equal(ops.dropout_add(x, r, 0.1, training=True), y)          # reproducible under manual_seed
    assert not torch.equal(ops.dropout_add(x, r, 0.1, training=True), y)      # and fresh on the next call
    m = (ops.dropout_add(torch.ones(8, 1 << 20, device=DEV), None, 0.5, True) != 0).float()
    assert abs(float(m.mean()) - 0.5) < 2e-3 and abs(float((m[:, 1:] * m[:, :-1]).mean()) - 0.25) < 2e-3
    yb = ops.dropout_add(x.bfloat16(), r, 0.0, training=True)                 # bf16 branch onto an fp32 stream
    assert yb.dtype == torch.float32 and rel_err(yb, x.bfloat16().float() + r) < 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gelu_dropout(dtype):
    g = torch.Generator().manual_seed(0)
    u = (torch.randn(64, 197, 3072, generator=g) * 1.5).to(dtype).float()
    cot = torch.randn(64, 197, 3072, generator=g).to(dtype).float()
    tol = TOL_F32 if dtype == torch.float32 else TOL_BF16
    ud = u.to(DEV, dtype).requires_grad_(True)
    out = ops.gelu_dropout(ud, 0.1, training=False)                          # eval: exact-erf GELU only
    out.backward(cot.to(DEV, dtype))
    uc = u.clone().requires_grad_(True)
    F.gelu(uc).backward(cot)
    assert rel_err(out, F.gelu(u)) < tol and rel_err(ud.grad, uc.grad) < tol
    torch.manual_seed(1)
    ud.grad = None
    out = ops.gelu_dropout(ud, 0.1, training=True)
    out.backward(cot.to(DEV, dtype))
    want_o, want_g = F.gelu(u).to(DEV) / 0.9, uc.grad.to(DEV) / 0.9
    atol_o, atol_g = tol * float(want_o.abs().max()), tol * float(want_g.abs().max())
    kept_o, kept_g = (out.float() - want_o).abs() <= atol_o, (ud.grad.float() - want_g).abs() <= atol_g
    assert (kept_o | (out == 0)).all() and (kept_g | (ud.grad == 0)).all()    # every element is either kept or dropped
    big = want_o.abs() > 10 * atol_o                                           # where kept / dropped is observable
    dropped = (out == 0) & big
    assert abs(float(dropped.float().sum() / big.float().sum()) - 0.1) < 5e-3  # Bernoulli(0.9) keep mask
    assert (ud.grad[dropped] == 0).all()                                        # the backward applies the same mask
    assert (kept_g | ~big | dropped).all()


def test_layernorm_fp32_stream_bf16_branch_under_autocast():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2 * 197, 768, generator=g) * 2
    w, b = 1 + 0.2 * torch.randn(768, generator=g), 0.1 * torch.randn(768, generator=g)
    cot = torch.randn(2 * 197, 768, generator=g).bfloat16().float()
    xd, wd, bd = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = ops.layer_norm(xd, wd, bd)
    assert y.dtype == torch.bfloat16                                          # the cast is folded into the kernel
    y.backward(cot.to(DEV, torch.bfloat16))
    assert xd.grad.dtype == torch.float32
    xc, wc, bc = (t.clone().requires_grad_(True) for t in (x, w, b))
    F.layer_norm(xc, (768,), wc, bc).backward(cot)
    assert rel_err(y, F.layer_norm(x, (768,), w, b)) < TOL_BF16
    assert rel_err(xd.grad, xc.grad) < TOL_F32 * 10 and rel_err(wd.grad, wc.grad) < TOL_F32 * 10
    y32 = ops.layer_norm(xd, wd, bd)                                           # no autocast: fp32 in, fp32 out
    assert y32.dtype == torch.float32 and rel_err(y32, F.layer_norm(x, (768,), w, b)) < TOL_F32


@pytest.mark.parametrize("rows,D,dtype", [(50432, 768, torch.bfloat16), (1000, 3072, torch.bfloat16), (7, 8, torch.float32),
                                           (12345, 2304, torch.bfloat16), (300, 72, torch.float32)])
def test_colsum_matches_fp64_sum(rows, D, dtype):
    g = torch.Generator(device=DEV).manual_seed(rows)
    x = torch.randn(rows, D, generator=g, device=DEV).to(dtype)
    got = ops.colsum(x)
    want = x.double().sum(0)
    assert got.dtype == torch.float32
    assert float((got.double() - want).abs().max()) <= 1e-4 * max(1.0, float(want.abs().max())) + 1e-3
    assert torch.equal(got, ops.colsum(x))                  # fixed reduction order
    if rows > 7:                                            # skip_period: rows 0, 7, 14, ... left out (CLS rows of a token tensor)
        keep = torch.ones(rows, dtype=torch.bool, device=x.device)
        keep[::7] = False
        want7 = x.double()[keep].sum(0)
        got7 = ops.colsum(x, skip_period=7)
        assert float((got7.double() - want7).abs().max()) <= 1e-5 * float(x.double().abs().sum(0).max()) + 1e-6
        xz = x.clone()
        xz[keep] = 0                                        # only skipped rows carry values: the sums are EXACTLY zero
        assert float(ops.colsum(xz, skip_period=7).abs().max()) == 0.0


def test_linear_matches_torch_linear_forward_and_backward():
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(4, 197, 256, generator=g, device=DEV, requires_grad=True)
    w = (torch.randn(512, 256, generator=g, device=DEV) * 0.05).requires_grad_(True)
    b = torch.randn(512, generator=g, device=DEV, requires_grad=True)
    cot = torch.randn(4, 197, 512, generator=g, device=DEV)
    y = ops.linear(x, w, b)
    y.backward(cot)
    got = (y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone())
    x.grad = w.grad = b.grad = None
    y2 = torch.nn.functional.linear(x, w, b)
    y2.backward(cot)
    for a, r in zip(got, (y2.detach(), x.grad, w.grad, b.grad)):
        assert rel_err(a, r) < 1e-5


@pytest.mark.parametrize("dtype,p", [(torch.float32, 0.0), (torch.float32, 0.3), (torch.bfloat16, 0.1)])
def test_fused_linear_edges_match_the_unfused_composition(dtype, p):
    """linear_gelu_dropout / linear_dropout_add (bias gradient from the edge's backward pass) against
    linear -> gelu_dropout / dropout_add with the same seeds."""
    g = torch.Generator(device=DEV).manual_seed(5)
    x0 = torch.randn(3, 197, 128, generator=g, device=DEV).to(dtype)
    r0 = torch.randn(3, 197, 128, generator=g, device=DEV).to(dtype)
    w1 = (torch.randn(256, 128, generator=g, device=DEV) * 0.1).to(dtype)
    b1 = torch.randn(256, generator=g, device=DEV).to(dtype)
    w2 = (torch.randn(128, 256, generator=g, device=DEV) * 0.1).to(dtype)
    b2 = torch.randn(128, generator=g, device=DEV).to(dtype)
    cot = torch.randn(3, 197, 128, generator=g, device=DEV).to(dtype)
    outs = []
    ops._FC1_ENABLED["on"] = False       # the fused fc1 GEMM draws a different (bit-sliced) mask: it has its own test below
    for fused in (True, False):
        leaves = [t.clone().requires_grad_(True) for t in (x0, r0, w1, b1, w2, b2)]
        x, r, W1, B1, W2, B2 = leaves
        torch.manual_seed(99)                                  # same dropout seeds in both variants
        if fused:
            h = ops.linear_gelu_dropout(x, W1, B1, p, True)
            y = ops.linear_dropout_add(h, W2, B2, r, p, True)
        else:
            h = ops.gelu_dropout(ops.linear(x, W1, B1), p, True)
            y = ops.dropout_add(ops.linear(h, W2, B2), r, p, True)
        y.backward(cot)
        outs.append([y.detach()] + [t.grad for t in leaves])
    ops._FC1_ENABLED["on"] = True
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for a, b in zip(*outs):
        assert rel_err(a, b) < tol


def _prologue_reference(img, w, b, cls, pos):
    """vit.py:34-36, 207-211 on the CPU in fp32."""
    y = F.conv2d(img, w, b, stride=w.shape[-1]).flatten(2).transpose(1, 2)
    return torch.cat((cls.expand(img.shape[0], -1, -1), y), dim=1) + pos


@pytest.mark.parametrize("B,C,S,P,D", [(3, 3, 224, 16, 768), (2, 3, 64, 8, 128), (1, 1, 48, 16, 64), (2, 3, 384, 16, 1024)])
@pytest.mark.parametrize("mode", ["fp32", "autocast"])
def test_patch_embed_tokens_forward_backward(B, C, S, P, D, mode):
    """f4 - PatchEmbed + CLS + pos_embed (vit.py:25-36, 207-211) as patchify + GEMM + assemble, against Conv2d."""
    g = torch.Generator().manual_seed(S + D)
    N = (S // P) ** 2 + 1
    img = torch.randn(B, C, S, S, generator=g)
    w = torch.randn(D, C, P, P, generator=g) * 0.05
    b, cls, pos = torch.randn(D, generator=g) * 0.1, torch.randn(1, 1, D, generator=g) * 0.1, torch.randn(1, N, D, generator=g) * 0.1
    cot = torch.randn(B, N, D, generator=g)
    ref_p = [t.clone().requires_grad_(True) for t in (w, b, cls, pos)]
    ref = _prologue_reference(img, *ref_p)
    ref.backward(cot)
    dev_p = [t.to(DEV).requires_grad_(True) for t in (w, b, cls, pos)]
    if mode == "autocast":
        ops.refresh_shadows(dev_p, torch.bfloat16)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = ops.patch_embed_tokens(img.to(DEV), *dev_p, 0.1, training=False)
        assert out.dtype == torch.bfloat16
        out.backward(cot.to(DEV, torch.bfloat16))
        tol = TOL_BF16
    else:
        out = ops.patch_embed_tokens(img.to(DEV), *dev_p, 0.1, training=False)
        out.backward(cot.to(DEV))
        tol = TOL_F32
    assert rel_err(out, ref) < tol
    for d, r, name in zip(dev_p, ref_p, ("weight", "bias", "cls", "pos")):
        assert d.grad.dtype == torch.float32 and rel_err(d.grad, r.grad) < tol, name


def test_patch_embed_tokens_dropout_and_errors():
    g = torch.Generator().manual_seed(5)
    img, w = torch.randn(4, 3, 64, 64, generator=g).to(DEV), (torch.randn(128, 3, 16, 16, generator=g) * 0.05).to(DEV)
    b, cls, pos = torch.zeros(128, device=DEV), torch.zeros(1, 1, 128, device=DEV), torch.zeros(1, 17, 128, device=DEV)
    pos.requires_grad_(True)
    clean = ops.patch_embed_tokens(img, w, b, cls, pos, 0.25, training=False)
    torch.manual_seed(3)
    y = ops.patch_embed_tokens(img, w, b, cls, pos, 0.25, training=True)
    y.backward(torch.ones_like(y))
    kept = y != 0
    assert abs(float(kept[:, 1:].float().mean()) - 0.75) < 2e-2
    assert rel_err(y[kept], clean[kept] / 0.75) < 1e-6
    # d pos = sum over the batch of keep / (1 - p)
    assert rel_err(pos.grad[0, 1:], kept[:, 1:].float().sum(0) / 0.75) < 1e-6
    torch.manual_seed(3)
    assert torch.equal(ops.patch_embed_tokens(img, w, b, cls, pos, 0.25, training=True), y)
    with pytest.raises(ValueError):
        ops.patch_embed_tokens(img, w[:, :, :12, :12].contiguous(), b, cls, pos, 0.0, False)     # 12 is not a multiple of 8
    with pytest.raises(RuntimeError):
        ops.patch_embed_tokens(img.cpu(), w, b, cls, pos, 0.0, False)


def test_parameter_shadows_follow_the_master():
    w = torch.nn.Parameter(torch.randn(64, 64, device=DEV))
    x = torch.randn(8, 64, device=DEV, dtype=torch.bfloat16)
    ops.refresh_shadows([w], torch.bfloat16)
    s0 = ops._shadow(w, torch.bfloat16)
    assert s0.dtype == torch.bfloat16 and torch.equal(s0, w.detach().bfloat16())
    assert ops._shadow(w, torch.bfloat16) is s0                        # fresh: reused, no cast
    with torch.no_grad():
        w.add_(1.0)                                                    # an optimizer step bumps the version
    s1 = ops._shadow(w, torch.bfloat16)
    assert s1 is not s0 and torch.equal(s1, w.detach().bfloat16())     # stale shadow is never served
    ops.refresh_shadows([w], torch.bfloat16)
    assert ops._shadow(w, torch.bfloat16) is s0 and torch.equal(s0, w.detach().bfloat16())    # refreshed in place
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = ops.linear(x, w, None)
    y.sum().backward()
    assert w.grad.dtype == torch.float32
    assert rel_err(w.grad, (torch.ones(8, 64, device=DEV).t() @ x.float())) < 1e-6          # fp32 accumulator written out unrounded


@pytest.mark.parametrize("M,N,K", [(4 * 197, 3072, 768), (130, 256, 64), (1, 512, 128), (50432, 3072, 768)])
def test_fused_fc1_gemm_epilogue(M, N, K):
    """f1 - fc1 + bias + GELU + dropout as ONE tcgen05 GEMM (gvit_linear_gelu_dropout_fwd) against a float64 GEMM of the
    same bf16 operands; the saved pre-activation, the keep mask and the backward through it."""
    assert ops.fused_fc1_available(N, K)
    g = torch.Generator(device=DEV).manual_seed(M + N)
    x = torch.randn(M, K, generator=g, device=DEV).bfloat16()
    w = (torch.randn(N, K, generator=g, device=DEV) * K ** -0.5).bfloat16()
    b = torch.randn(N, generator=g, device=DEV).bfloat16()
    rows = torch.randperm(M, generator=torch.Generator().manual_seed(1))[:256].to(DEV)        # sample rows for the fp64 check
    u_ref = (x[rows].double() @ w.double().t() + b.double()).bfloat16()
    want = torch.nn.functional.gelu(u_ref.double())
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    out = ops.linear_gelu_dropout(xr, wr, br, 0.0, True)
    assert out.shape == (M, N) and out.dtype == torch.bfloat16
    # u is rounded to bf16 before the GELU (as the unfused path stores it): one bf16 ulp of u may flip
    assert rel_err(out[rows], want) < TOL_BF16
    assert float((out[rows].detach().double() - want).abs().mean() / want.abs().mean()) < 2e-3
    torch.manual_seed(7)
    outp = ops.linear_gelu_dropout(xr, wr, br, 0.25, True)
    kept = outp != 0
    dead = out == 0                                             # exact zeros of the GELU itself (u = -0 ... underflow)
    assert abs(float(kept[~dead].float().mean()) - 0.75) < max(5e-3, 4.0 * (0.1875 / max(1, int((~dead).sum()))) ** 0.5)
    assert rel_err(outp[kept].float(), (out.float() / 0.75)[kept]) < 1e-2
    torch.manual_seed(7)
    assert torch.equal(ops.linear_gelu_dropout(xr, wr, br, 0.25, True), outp)                # reproducible
    if M <= 1024:
        cot = torch.randn(M, N, generator=g, device=DEV).bfloat16()
        outp.backward(cot)
        ops._FC1_ENABLED["on"] = False                          # the unfused composition with the same seed
        try:
            x2, w2, b2 = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
            torch.manual_seed(7)
            ref = ops.linear_gelu_dropout(x2, w2, b2, 0.25, True)
            ref.backward(cot)
        finally:
            ops._FC1_ENABLED["on"] = True
        # the fused epilogue draws its keep bits bit-sliced (a different, equally valid assignment of the same Philox
        # stream), so the two masks differ: compare where both kept, and the gradients through each path's OWN mask
        both = kept & (ref != 0)
        assert abs(float(both.float().mean()) - 0.75 ** 2) < max(1e-2, 4.0 / (M * N) ** 0.5) or bool(dead.float().mean() > 0.01)
        assert rel_err(outp[both].float(), ref[both].float()) < TOL_BF16
        m1, m2 = kept.float() / 0.75, (ref != 0).float() / 0.75
        gu = torch.nn.functional.gelu((x.float() @ w.float().t() + b.float()).bfloat16().float().requires_grad_(True))
        # d out / d x for a given mask, by autograd on the unfused formula in fp32
        for mk, xg, wg, bg in ((m1, xr.grad, wr.grad, br.grad), (m2, x2.grad, w2.grad, b2.grad)):
            x3, w3, b3 = x.float().requires_grad_(True), w.float().requires_grad_(True), b.float().requires_grad_(True)
            (torch.nn.functional.gelu(x3 @ w3.t() + b3) * mk).backward(cot.float())
            assert rel_err(xg, x3.grad) < TOL_BF16 and rel_err(wg, w3.grad) < TOL_BF16 and rel_err(bg, b3.grad) < TOL_BF16


@pytest.mark.parametrize("M,N,K", [(4 * 197, 768, 768), (130, 256, 64), (50432, 768, 768)])
def test_fused_proj_gemm_epilogue(M, N, K):
    """f1 - proj + proj_drop + residual add as ONE tcgen05 GEMM (gvit_linear_dropout_residual_fwd) against a float64 GEMM
    of the same bf16 operands; keep mask statistics; the backward through the kernel's own mask."""
    assert ops.fused_fc1_available(N, K) and K <= ops._FUSED_RESID_MAX_K
    g = torch.Generator(device=DEV).manual_seed(M + N + 1)
    x = torch.randn(M, K, generator=g, device=DEV).bfloat16()
    w = (torch.randn(N, K, generator=g, device=DEV) * K ** -0.5).bfloat16()
    b = torch.randn(N, generator=g, device=DEV).bfloat16()
    r = torch.randn(M, N, generator=g, device=DEV).bfloat16()
    rows = torch.randperm(M, generator=torch.Generator().manual_seed(1))[:256].to(DEV)
    y_ref = (x[rows].double() @ w.double().t() + b.double()).bfloat16().double()
    before = ops.launch_count()
    out0 = ops.linear_dropout_add(x, w, b, r, 0.0, True)
    assert ops.launch_count() == before + 1                    # one kernel: no separate GEMM + edge pass
    assert rel_err(out0[rows], r[rows].double() + y_ref) < TOL_BF16
    xr, wr, br, rr = (t.clone().requires_grad_(True) for t in (x, w, b, r))
    torch.manual_seed(11)
    out = ops.linear_dropout_add(xr, wr, br, rr, 0.25, True)
    y_all = (out0.float() - r.float())                        # x W^T + b as the kernel computed it (up to bf16 rounding of out0)
    kept = (out.float() - r.float()).abs() > 0.5 * y_all.abs() / 0.75
    live = y_all.abs() > 0.05                                  # keep decisions are only visible where y is not ~0
    assert abs(float(kept[live].float().mean()) - 0.75) < max(5e-3, 4.0 * (0.1875 / max(1, int(live.sum()))) ** 0.5)
    torch.manual_seed(11)
    assert torch.equal(ops.linear_dropout_add(xr, wr, br, rr, 0.25, True), out)             # reproducible
    if M <= 1024:
        # backward through the kernel's OWN mask: with a zero residual the mask is readable from the output (out = keep * y / (1-p))
        cot = torch.randn(M, N, generator=g, device=DEV).bfloat16()
        xz, wz, bz, rz = (t.clone().requires_grad_(True) for t in (x, w, b, torch.zeros_like(r)))
        torch.manual_seed(12)
        oz = ops.linear_dropout_add(xz, wz, bz, rz, 0.25, True)
        oz.backward(cot)
        mask = (oz != 0).float()
        x3, w3, b3, r3 = (t.float().requires_grad_(True) for t in (x, w, b, torch.zeros_like(r)))
        (r3 + (x3 @ w3.t() + b3) * mask / 0.75).backward(cot.float())
        assert rel_err(rz.grad, r3.grad) < TOL_BF16 and rel_err(xz.grad, x3.grad) < TOL_BF16
        assert rel_err(wz.grad, w3.grad) < TOL_BF16 and rel_err(bz.grad, b3.grad) < TOL_BF16


@pytest.mark.parametrize("M,N,K,p", [(4 * 197, 3072, 768, 0.1), (130, 256, 64, 0.25), (300, 512, 128, 0.0), (50432, 3072, 768, 0.1)])
def test_fused_fc2_dgrad_gelu_backward_kernel(M, N, K, p):
    """f1 - gvit_linear_gelu_dropout_bwd (fc2 input-gradient GEMM + keep mask + GELU' + fc1 bias gradient in one kernel, W2
    read as stored) against the unfused pair: library GEMM, then gvit_gelu_dropout_bwd with the SAME mask bytes."""
    from graph_augmented_vision_transformers_b200 import _lib
    from graph_augmented_vision_transformers_b200.ops import _call, _ptr, _stream, GVIT_BF16, _colsum_ws
    g = torch.Generator(device=DEV).manual_seed(M + N + 3)
    dy = torch.randn(M, K, generator=g, device=DEV).bfloat16()                 # gradient of fc2's output
    w2 = (torch.randn(K, N, generator=g, device=DEV) * K ** -0.5).bfloat16()   # fc2.weight: (out = K here, in = N)
    u = (torch.randn(M, N, generator=g, device=DEV) * 1.5).bfloat16()
    mask = torch.randint(0, 256, (M * N // 8,), generator=g, device=DEV, dtype=torch.uint8) if p > 0 else None
    du = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    db = torch.empty(N, device=DEV)
    rows = _lib.load().gvit_linear_gelu_dropout_bwd_ws_rows(M)
    part = torch.empty(rows * N, device=DEV)
    _call("gvit_linear_gelu_dropout_bwd", _ptr(dy), _ptr(w2), _ptr(u), _ptr(mask), M, N, K, float(p), GVIT_BF16, 0, _ptr(du), _ptr(db),
          _ptr(part), _stream())
    dh = dy @ w2                                                               # (M, N) bf16, what the unfused path stores
    du_ref = torch.empty_like(du)
    db_ref = torch.empty(N, device=DEV)
    ws = _colsum_ws(M, N, DEV)
    _call("gvit_gelu_dropout_bwd", _ptr(dh), _ptr(u), _ptr(mask), M * N, float(p), GVIT_BF16, _ptr(du_ref), N, _ptr(db_ref), _ptr(ws),
          _stream())
    assert rel_err(du, du_ref) < TOL_BF16                                      # dh is not rounded to bf16 on the fused path
    assert float((du.float() - du_ref.float()).abs().mean() / du_ref.float().abs().mean()) < 4e-3
    assert ((du == 0) == (du_ref == 0)).float().mean() > 0.999                 # same mask applied
    assert rel_err(db, du.float().sum(0)) < 1e-4                               # the bias gradient sums the values as stored
    assert rel_err(db, db_ref) < TOL_BF16
    db2 = torch.empty(N, device=DEV)
    _call("gvit_linear_gelu_dropout_bwd", _ptr(dy), _ptr(w2), _ptr(u), _ptr(mask), M, N, K, float(p), GVIT_BF16, 0, _ptr(du_ref), _ptr(db2),
          _ptr(part), _stream())
    assert torch.equal(db, db2) and torch.equal(du, du_ref)                    # deterministic


@pytest.mark.parametrize("M,N,K,p", [(4 * 197, 3072, 768, 0.1), (130, 256, 64, 0.25), (300, 512, 128, 0.0)])
def test_fused_fc1_saved_backward_factor(M, N, K, p):
    """f1 - save_mode 1 of gvit_linear_gelu_dropout_fwd: `u` receives keep * gelu'(u) / (1 - p) instead of the pre-activation
    (same activation output, same keep decisions as save_mode 0), and gvit_linear_gelu_dropout_bwd(saved_mode 1) - one
    multiply in its epilogue - gives the gradient of the save_mode 0 pair; u == NULL saves nothing."""
    from graph_augmented_vision_transformers_b200 import _lib
    from graph_augmented_vision_transformers_b200.ops import _call, _ptr, _stream, GVIT_BF16
    g = torch.Generator(device=DEV).manual_seed(M + N + 11)
    x = torch.randn(M, K, generator=g, device=DEV).bfloat16()
    w = (torch.randn(N, K, generator=g, device=DEV) * K ** -0.5).bfloat16()
    b = torch.randn(N, generator=g, device=DEV).bfloat16()
    u0, o0, f1, o1, o2 = (torch.empty(M, N, device=DEV, dtype=torch.bfloat16) for _ in range(5))
    mask = torch.empty(M * N // 8, dtype=torch.uint8, device=DEV) if p > 0 else None
    st = _stream()
    _call("gvit_linear_gelu_dropout_fwd", _ptr(x), _ptr(w), _ptr(b), M, N, K, float(p), 77, 0, None, GVIT_BF16, 0, _ptr(u0), _ptr(o0), _ptr(mask), st)
    _call("gvit_linear_gelu_dropout_fwd", _ptr(x), _ptr(w), _ptr(b), M, N, K, float(p), 77, 0, None, GVIT_BF16, 1, _ptr(f1), _ptr(o1), None, st)
    _call("gvit_linear_gelu_dropout_fwd", _ptr(x), _ptr(w), _ptr(b), M, N, K, float(p), 77, 0, None, GVIT_BF16, 1, None, _ptr(o2), None, st)
    assert torch.equal(o0, o1) and torch.equal(o0, o2)                         # the activation does not depend on what is saved
    uf = u0.float()
    gprime = 0.5 * (1 + torch.erf(uf / 2 ** 0.5)) + uf * torch.exp(-0.5 * uf * uf) / (2 * torch.pi) ** 0.5
    if p > 0:
        bits = torch.from_numpy(__import__("numpy").unpackbits(mask.cpu().numpy(), bitorder="little")).to(DEV).view(M, N).float()
    else:
        bits = torch.ones(M, N, device=DEV)
    want = gprime * bits / (1 - p)
    assert rel_err(f1, want) < TOL_BF16 / 2
    assert ((f1 == 0) == (want.bfloat16() == 0)).float().mean() > 0.999
    # backward: (dy W2) * factor against the save_mode 0 kernel on (u, mask)
    dy = torch.randn(M, K, generator=g, device=DEV).bfloat16()
    w2 = (torch.randn(K, N, generator=g, device=DEV) * K ** -0.5).bfloat16()
    rows = _lib.load().gvit_linear_gelu_dropout_bwd_ws_rows(M)
    part = torch.empty(rows * N, device=DEV)
    du0, du1 = torch.empty_like(u0), torch.empty_like(u0)
    db0, db1 = torch.empty(N, device=DEV), torch.empty(N, device=DEV)
    _call("gvit_linear_gelu_dropout_bwd", _ptr(dy), _ptr(w2), _ptr(u0), _ptr(mask), M, N, K, float(p), GVIT_BF16, 0, _ptr(du0), _ptr(db0), _ptr(part), st)
    _call("gvit_linear_gelu_dropout_bwd", _ptr(dy), _ptr(w2), _ptr(f1), None, M, N, K, float(p), GVIT_BF16, 1, _ptr(du1), _ptr(db1), _ptr(part), st)
    assert rel_err(du1, du0) < TOL_BF16                                        # the factor is rounded to bf16 once more
    assert float((du1.float() - du0.float()).abs().mean() / du0.float().abs().mean()) < 6e-3
    assert rel_err(db1, du1.float().sum(0)) < 1e-4 and rel_err(db1, db0) < TOL_BF16


@pytest.mark.parametrize("p", [0.0, 0.2])
def test_mlp_fused_node_matches_the_two_op_composition(p):
    """The whole-Mlp autograd node (fused fc1 forward, fused fc2-dgrad backward) against linear_gelu_dropout +
    linear_dropout_add on the library GEMMs; with dropout the two draw different masks, so p > 0 compares statistics."""
    g = torch.Generator(device=DEV).manual_seed(9)
    x0 = torch.randn(3, 197, 128, generator=g, device=DEV).bfloat16()
    r0 = torch.randn(3, 197, 128, generator=g, device=DEV).bfloat16()
    w1 = (torch.randn(512, 128, generator=g, device=DEV) * 0.1).bfloat16()
    b1 = (torch.randn(512, generator=g, device=DEV) * 0.1).bfloat16()
    w2 = (torch.randn(128, 512, generator=g, device=DEV) * 0.05).bfloat16()
    b2 = (torch.randn(128, generator=g, device=DEV) * 0.1).bfloat16()
    cot = torch.randn(3, 197, 128, generator=g, device=DEV).bfloat16()
    assert ops.mlp_fused_available(x0, w1, w2, r0)
    outs = []
    for fused in (True, False):
        leaves = [t.clone().requires_grad_(True) for t in (x0, w1, b1, w2, b2, r0)]
        x, W1, B1, W2, B2, r = leaves
        torch.manual_seed(4)
        if fused:
            y = ops.mlp_fused(x, W1, B1, W2, B2, r, p, True)
        else:
            ops._FC1_ENABLED["on"] = False
            try:
                y = ops.linear_dropout_add(ops.linear_gelu_dropout(x, W1, B1, p, True), W2, B2, r, p, True)
            finally:
                ops._FC1_ENABLED["on"] = True
        y.backward(cot)
        outs.append([y.detach()] + [t.grad for t in leaves])
    if p == 0.0:
        for a, b in zip(*outs):
            assert rel_err(a, b) < TOL_BF16
    else:
        for a, b in zip(*outs):                               # different masks: same scale of every quantity
            assert torch.isfinite(a.float()).all() and 0.8 < float(a.float().norm() / b.float().norm()) < 1.25
        assert torch.equal(outs[0][6], cot)                   # d resid is the incoming gradient itself
