"""a7+a8 - the graph sub-layer forward and backward through gvit_knn_fwd, gvit_agg_fwd /
gvit_agg_gather_fwd, gvit_graph_reverse, gvit_agg_bwd, gvit_knn_bwd, against the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import TOL_BF16, TOL_F32, golden, rel_err
from gpu_util import DEV, check_adjacency, tokens
from graph_augmented_vision_transformers_b200 import modules, ops
from oracle import graph_oracle

pytestmark = pytest.mark.gpu


def _run_device(hc, W, b, k, dtype, cot):
    h = hc.to(DEV, dtype).requires_grad_(True)
    Wd = W.to(DEV, dtype).requires_grad_(True)
    bd = b.to(DEV, dtype).requires_grad_(True)
    out, idx, vals = ops.patch_graph(h, Wd, bd, k, return_graph=True)
    out.backward(cot.to(DEV, dtype))
    return out, idx, vals, h.grad, Wd.grad, bd.grad


def _run_oracle(hc, W, b, k, cot, idx, compute_dtype=None):
    h = hc.clone().requires_grad_(True)
    Wc = W.clone().requires_grad_(True)
    bc = b.clone().requires_grad_(True)
    # idx None: the oracle builds its own (strict fp32) adjacency - no help from the device result
    out, aux = graph_oracle.graph_layer_forward(h, Wc, bc, k, "knn", compute_dtype=compute_dtype,
                                                idx_override=None if idx is None else idx.cpu(), return_aux=True)
    out.float().backward(cot)
    return out.float(), h.grad, Wc.grad, bc.grad, aux["idx"]


@pytest.mark.parametrize("name", ["graph_knn_small", "graph_knn_196"])
def test_fp32_golden_forward_backward(name):
    g = golden(name)
    hc, W, b, cot = (torch.from_numpy(g[n]) for n in ("h", "W", "b", "cot"))
    out, idx, vals, dh, dW, db = _run_device(hc, W, b, int(g["k"]), torch.float32, cot)
    assert np.array_equal(idx.cpu().numpy(), g["idx"]) and np.array_equal(vals.cpu().numpy(), g["vals"])   # strict fp32, all rows
    assert rel_err(out, g["out"]) < TOL_F32
    assert rel_err(dh, g["dh"]) < TOL_F32 and rel_err(dW, g["dW"]) < TOL_F32 and rel_err(db, g["db"]) < TOL_F32
    want = _run_oracle(hc, W, b, int(g["k"]), cot, None)
    assert torch.equal(want[4], idx.cpu().long())
    for got, ref, n in zip((out, dh, dW, db), want, ("out", "dh", "dW", "db")):
        assert rel_err(got, ref) < TOL_F32, n
    assert float(out.detach()[:, 0].abs().max()) == 0.0 and float(dh[:, 0].abs().max()) == 0.0


@pytest.mark.parametrize("B,Np,D,k", [(2, 196, 768, 8), (1, 50, 72, 5), (2, 300, 128, 16)])
def test_fp32_forward_backward_vs_oracle(B, Np, D, k):
    hc, _ = tokens(B, Np, D, seed=11)
    g = torch.Generator().manual_seed(3)
    W, b, cot = torch.randn(D, D, generator=g) * 0.05, torch.randn(D, generator=g) * 0.1, torch.randn(B, Np + 1, D, generator=g)
    out, idx, vals, dh, dW, db = _run_device(hc, W, b, k, torch.float32, cot)
    check_adjacency(hc, idx, vals, k, noise=2e-6)
    want = _run_oracle(hc, W, b, k, cot, None)               # the oracle's own strict-order adjacency
    assert torch.equal(want[4], idx.cpu().long())            # bit-exact indices on every row
    for got, ref, n in zip((out, dh, dW, db), want, ("out", "dh", "dW", "db")):
        assert rel_err(got, ref) < TOL_F32, n


@pytest.mark.parametrize("B,Np,D,k", [(3, 196, 768, 8), (2, 196, 768, 4), (2, 196, 768, 16), (2, 64, 128, 8),
                                      (1, 16, 64, 2), (2, 256, 1024, 8), (2, 129, 192, 8), (2, 196, 768, 32),
                                      (1, 576, 1024, 8),
                                      # the CTA-pair kernels off the bench shape: half-slab split (NT = 144 / 160 / 256), odd and
                                      # even chunk counts (D = 128 ... 640: split and unsplit work items), k that is not a vector row
                                      (3, 129, 128, 8), (2, 144, 384, 8), (5, 256, 256, 4), (3, 150, 640, 5), (2, 200, 512, 16),
                                      (1, 256, 768, 8)])
def test_bf16_forward_backward_vs_oracle(B, Np, D, k):
    """bf16 (tcgen05 kernels where the shape is in range): oracle = autocast semantics, G1-G4 fp32, G5-G6 on bf16."""
    bf = torch.bfloat16
    hc, _ = tokens(B, Np, D, seed=21, dtype=bf)
    g = torch.Generator().manual_seed(4)
    W = (torch.randn(D, D, generator=g) * 0.05).to(bf).float()
    b = (torch.randn(D, generator=g) * 0.1).to(bf).float()
    cot = torch.randn(B, Np + 1, D, generator=g).to(bf).float()
    out, idx, vals, dh, dW, db = _run_device(hc, W, b, k, bf, cot)
    assert out.dtype == bf and dh.dtype == bf
    check_adjacency(hc, idx, vals, k, noise=1e-5)
    want = _run_oracle(hc, W, b, k, cot, idx, compute_dtype=bf)
    for got, ref, n in zip((out, dh, dW, db), want, ("out", "dh", "dW", "db")):
        assert rel_err(got, ref) < TOL_BF16, n
    assert float(out.detach()[:, 0].abs().max()) == 0.0 and float(dh[:, 0].abs().max()) == 0.0


def test_bf16_fused_residual_epilogue():
    bf = torch.bfloat16
    hc, hd = tokens(2, 196, 768, seed=2, dtype=bf)
    g = torch.Generator().manual_seed(5)
    W = (torch.randn(768, 768, generator=g) * 0.05).to(DEV, bf)
    b = (torch.randn(768, generator=g) * 0.1).to(DEV, bf)
    x = torch.randn(2, 197, 768, generator=g).to(DEV, bf).requires_grad_(True)
    plain = ops.patch_graph(hd, W, b, 8)
    fused = ops.patch_graph(hd, W, b, 8, resid=x)
    assert rel_err(fused, x.float() + plain.float()) < 1e-2
    assert torch.equal(fused[:, 0], x[:, 0])                  # CLS row passes through untouched (G0)
    fused.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
    x32 = x.detach().float()                                  # fp32 residual stream under autocast
    with torch.autocast("cuda", dtype=bf):
        mixed = ops.patch_graph(hd.float(), W.float(), b.float(), 8, resid=x32)
    assert mixed.dtype == torch.float32 and rel_err(mixed, x32 + plain.float()) < 1e-6


@pytest.mark.parametrize("B", [1, 37, 75, 80, 120])
@pytest.mark.parametrize("res32", [False, True])
def test_bf16_aggregation_work_items_are_batch_invariant(B, res32):
    """The CTA-pair aggregation kernel schedules whole images in full rounds and splits the left-over images' output chunks
    between two pairs (agg4_tc.cu get_item): every image must come out bit-identical to the same image run in a batch of one
    or two, whatever round / half it landed in - output, CLS row, and the saved tiles the backward reads."""
    bf = torch.bfloat16
    Np, D, k = 196, 768, 8
    hc, hd = tokens(B, Np, D, seed=40 + B, dtype=bf)
    g = torch.Generator().manual_seed(6)
    W = (torch.randn(D, D, generator=g) * 0.05).to(DEV, bf)
    b = (torch.randn(D, generator=g) * 0.1).to(DEV, bf)
    x = torch.randn(B, Np + 1, D, generator=g).to(DEV, torch.float32 if res32 else bf)
    cot = torch.randn(B, Np + 1, D, generator=g).to(DEV, x.dtype)

    def run(h, xr, c):
        h = h.clone().requires_grad_(True)
        if res32:
            with torch.autocast("cuda", dtype=bf):
                out = ops.patch_graph(h.float(), W.float(), b.float(), k, resid=xr)
        else:
            out = ops.patch_graph(h, W, b, k, resid=xr)
        out.backward(c)
        return out.detach(), h.grad

    out, dh = run(hd, x, cot)
    for lo in sorted({0, B // 2, max(B - 2, 0)}):
        hi = min(lo + 2, B)
        o2, d2 = run(hd[lo:hi], x[lo:hi], cot[lo:hi])
        assert torch.equal(out[lo:hi], o2), (lo, "out")
        assert torch.equal(dh[lo:hi], d2), (lo, "dh")


def test_bf16_aggregation_skips_out_of_range_neighbours():
    """The adjacency lists are caller-supplied pointers at the C ABI: an index outside [0, Np) must not become a shared-memory
    scatter.  The product build drops such an edge (a -DGVIT_DEBUG_BOUNDS build traps with a message): every other row of the
    image, and every other image, comes out exactly as with the clean list."""
    from graph_augmented_vision_transformers_b200.ops import _call, _dtype_code, _ptr, _stream
    bf = torch.bfloat16
    B, Np, D, k = 3, 196, 768, 8
    _, hd = tokens(B, Np, D, seed=9, dtype=bf)
    g = torch.Generator().manual_seed(8)
    W = (torch.randn(D, D, generator=g) * 0.05).to(DEV, bf)
    b = (torch.randn(D, generator=g) * 0.1).to(DEV, bf)
    idx, vals, _ = ops.knn_graph(hd, k)

    def run(ix):
        out = torch.empty_like(hd)
        w = torch.empty(B, Np, k, device=DEV)
        z = torch.empty(B, Np, D, device=DEV, dtype=bf)
        _call("gvit_agg_fwd", _ptr(hd), B, Np, D, k, _dtype_code(hd), _ptr(ix), _ptr(vals), _ptr(W), _ptr(b), None, _dtype_code(hd),
              _ptr(out), _ptr(w), _ptr(z), Np * D, _stream())
        torch.cuda.synchronize()
        return out

    clean = run(idx)
    bad = idx.clone()
    bad[1, 7, 3] = 1_000_000
    bad[1, 150, 0] = -5
    got = run(bad)
    keep = torch.ones(B, Np + 1, dtype=torch.bool, device=DEV)
    keep[1, 1 + 7] = False
    keep[1, 1 + 150] = False
    assert torch.equal(got[keep], clean[keep])
    assert torch.isfinite(got.float()).all()


def test_graph_reverse_is_the_transposed_adjacency():
    _, hd = tokens(3, 196, 64, seed=9)
    idx, _, _ = ops.knn_graph(hd, 8)
    rev_ptr, rev_src = ops.graph_reverse(idx)
    idx, rev_ptr, rev_src = idx.cpu().numpy(), rev_ptr.cpu().numpy(), rev_src.cpu().numpy()
    for b in range(3):
        flat = idx[b].reshape(-1)
        assert rev_ptr[b, 0] == 0 and rev_ptr[b, -1] == flat.size
        for j in (0, 1, 57, 195):
            edges = rev_src[b, rev_ptr[b, j]:rev_ptr[b, j + 1]]
            assert np.array_equal(edges, np.nonzero(flat == j)[0])      # ascending edge ids -> fixed summation order


def test_dense_mode_vs_oracle():
    hc, hd = tokens(2, 100, 64, seed=6)
    layer = modules.PatchGraphLayer(64, mode="dense").to(DEV)
    want = graph_oracle.graph_layer_forward(hc, layer.proj.weight.detach().cpu(), layer.proj.bias.detach().cpu(), 0, "dense")
    assert rel_err(layer(hd), want) < TOL_F32


@pytest.mark.parametrize("batch,M,N,K,a_t,b_t,f32", [(3, 196, 196, 768, 0, 0, True), (2, 576, 1024, 576, 0, 1, False),
                                                       (2, 576, 1024, 576, 1, 1, False), (1, 100, 64, 128, 0, 0, False),
                                                       (2, 130, 192, 64, 1, 1, True), (5, 576, 576, 1024, 0, 0, True),
                                                       (1, 64, 320, 192, 0, 1, False)])
def test_bgemm_all_operand_majors_vs_torch(batch, M, N, K, a_t, b_t, f32):
    """gvit_bgemm (tcgen05, persistent, operand majors by descriptor) against torch.bmm in fp32 on the same bf16 values;
    operands live inside padded rows the way the layer's tensors do."""
    from graph_augmented_vision_transformers_b200.ops import _bgemm, _ptr
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    bf = torch.bfloat16
    pad = lambda n: (n + 63) // 64 * 64

    def operand(rows, cols):                    # (batch, rows, pad(cols)) storage, zero padding
        t = torch.zeros(batch, rows, pad(cols), device=DEV, dtype=bf)
        t[:, :, :cols] = torch.randn(batch, rows, cols, generator=g, device=DEV).to(bf)
        return t

    A = operand(K, M) if a_t else operand(M, K)
    Bm = operand(K, N) if b_t else operand(N, K)
    Al = (A[:, :, :M].transpose(1, 2) if a_t else A[:, :, :K]).float()            # logical (M, K)
    Bl = (Bm[:, :, :N] if b_t else Bm[:, :, :K].transpose(1, 2)).float()          # logical (K, N)
    rs = torch.rand(batch, M, generator=g, device=DEV) + 0.5
    out = torch.full((batch, M, pad(N)), 7.0, device=DEV, dtype=torch.float32 if f32 else bf)
    _bgemm(batch, M, N, [(_ptr(A), A.shape[2], A.shape[1] * A.shape[2], a_t, _ptr(Bm), Bm.shape[2], Bm.shape[1] * Bm.shape[2], b_t, K)],
           out, out.shape[2], out.shape[1] * out.shape[2], row_scale=rs)
    want = torch.bmm(Al, Bl) * rs[:, :, None]
    assert rel_err(out[:, :, :N], want) < (1e-5 if f32 else 1e-2)
    per = 4 if f32 else 8                                                 # rows are written in whole 16-byte chunks
    Nc = (N + per - 1) // per * per
    assert float((out[:, :, Nc:].float() - 7.0).abs().max() if pad(N) > Nc else 0.0) == 0.0   # nothing written beyond them
    if a_t == 0 and K == M and not f32:          # two products into one accumulator: X Y + X^T Y
        out2 = torch.empty_like(out)
        _bgemm(batch, M, N, [(_ptr(A), A.shape[2], A.shape[1] * A.shape[2], 0, _ptr(Bm), Bm.shape[2], Bm.shape[1] * Bm.shape[2], b_t, K),
                             (_ptr(A), A.shape[2], A.shape[1] * A.shape[2], 1, _ptr(Bm), Bm.shape[2], Bm.shape[1] * Bm.shape[2], b_t, K)],
               out2, out.shape[2], out.shape[1] * out.shape[2])
        assert rel_err(out2[:, :, :N], torch.bmm(Al + Al.transpose(1, 2), Bl)) < 1e-2


@pytest.mark.parametrize("B,Np,D", [(2, 196, 768), (1, 576, 1024), (3, 100, 64), (2, 129, 192)])
@pytest.mark.parametrize("res32", [False, True])
def test_dense_mode_bf16_native_forward_backward_vs_oracle(B, Np, D, res32):
    """graph_mode='dense' under bf16 compute: batched tcgen05 GEMMs + row-wise libgvit kernels (ops._DenseGraph) against the
    section-9 oracle with autocast semantics (G1-G4 fp32, G5-G6 on bf16), forward and every gradient."""
    bf = torch.bfloat16
    assert ops.dense_graph_available(bf, Np, D)
    hc, _ = tokens(B, Np, D, seed=31, dtype=bf)
    g = torch.Generator().manual_seed(6)
    W = (torch.randn(D, D, generator=g) * 0.05).to(bf).float()
    b = (torch.randn(D, generator=g) * 0.1).to(bf).float()
    x = torch.randn(B, Np + 1, D, generator=g).to(bf).float()
    cot = torch.randn(B, Np + 1, D, generator=g).to(bf).float()
    h = hc.to(DEV, bf).requires_grad_(True)
    Wd, bd = W.to(DEV, bf).requires_grad_(True), b.to(DEV, bf).requires_grad_(True)
    xd = x.to(DEV, torch.float32 if res32 else bf).requires_grad_(True)
    out = ops.dense_graph(h, Wd, bd, resid=xd)
    assert out.dtype == xd.dtype
    out.backward(cot.to(DEV, out.dtype))
    ho, Wo, bo = hc.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    want = graph_oracle.graph_layer_forward(ho, Wo, bo, 0, "dense", compute_dtype=bf).float()
    want.backward(cot)
    if res32:                                                            # fp32 stream: the branch can be isolated exactly
        assert rel_err(out.float() - xd.detach().float(), want) < TOL_BF16
    plain = ops.dense_graph(h.detach(), Wd.detach(), bd.detach())        # no residual: the branch itself, bf16
    assert plain.dtype == bf and rel_err(plain, want) < TOL_BF16 and float(plain[:, 0].abs().max()) == 0.0
    assert rel_err(out, x + want) < TOL_BF16
    assert torch.equal(out[:, 0], xd.detach()[:, 0])                     # CLS row passes through untouched (G0)
    assert torch.equal(xd.grad.float(), cot.to(DEV))                     # residual path: identity
    for got, ref, n in ((h.grad, ho.grad, "dh"), (Wd.grad, Wo.grad, "dW"), (bd.grad, bo.grad, "db")):
        assert rel_err(got, ref) < TOL_BF16, n
    assert float(h.grad[:, 0].abs().max()) == 0.0


def test_backward_is_deterministic_at_full_size():
    bf = torch.bfloat16
    g = torch.Generator(device=DEV).manual_seed(8)
    h = torch.randn(256, 197, 768, generator=g, device=DEV, dtype=bf)
    W = torch.randn(768, 768, generator=g, device=DEV, dtype=bf) * 0.05
    outs = []
    for _ in range(2):
        hh = h.clone().requires_grad_(True)
        o = ops.patch_graph(hh, W, None, 8)
        o.backward(torch.ones_like(o))
        outs.append((o.detach(), hh.grad))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])   # no atomics anywhere
    assert torch.isfinite(outs[0][0].float()).all() and torch.isfinite(outs[0][1].float()).all()
    # linearity in the projection: scaling Wg scales the output
    o2 = ops.patch_graph(h, W * 2, None, 8)
    assert rel_err(o2, 2 * outs[0][0].float()) < 1e-2


@pytest.mark.parametrize("B,Np,D,k", [(3, 196, 768, 8), (2, 144, 256, 4), (5, 64, 128, 16), (2, 129, 192, 8), (1, 16, 64, 2)])
def test_bf16_fused_backward_matches_gather_backward(B, Np, D, k):
    """gvit_graph_bwd (two per-image tensor-core GEMMs) against gvit_graph_reverse + gvit_agg_bwd + gvit_knn_bwd
    (reverse-CSR gathers) on the same saved tensors: same algebra, only bf16 rounding of the coefficients differs."""
    from graph_augmented_vision_transformers_b200.ops import _call, _dtype_code, _ptr, _stream, _token_view
    bf = torch.bfloat16
    _, hd = tokens(B, Np, D, seed=31, dtype=bf)
    assert ops.fused_graph_bwd_available(bf, Np, D, k)
    idx, vals, rnorm = ops.knn_graph(hd, k)
    w, _ = ops.agg_gather(hd, idx, vals)
    g = torch.Generator(device=DEV).manual_seed(7)
    dz = torch.randn(B, Np, D, generator=g, device=DEV).to(bf)
    off, bs, rs, _, _, _ = _token_view(hd)
    dt, st = _dtype_code(hd), _stream()
    rev_ptr, rev_src = ops.graph_reverse(idx)
    dvals_ref = torch.empty(B, Np, k, device=DEV)
    dh_ref = torch.zeros_like(hd)
    _call("gvit_agg_bwd", _ptr(hd, off), bs, rs, B, Np, D, k, dt, _ptr(idx), _ptr(w), _ptr(dz), _ptr(rev_ptr), _ptr(rev_src),
          _ptr(dvals_ref), _ptr(dh_ref, off), st)
    _call("gvit_knn_bwd", _ptr(hd, off), bs, rs, B, Np, D, k, dt, _ptr(idx), _ptr(rnorm), _ptr(dvals_ref), _ptr(rev_ptr),
          _ptr(rev_src), _ptr(dh_ref, off), st)
    dvals = torch.empty(B, Np, k, device=DEV)
    dh = torch.zeros_like(hd)
    _call("gvit_graph_bwd", _ptr(hd, off), bs, rs, B, Np, D, k, dt, _ptr(idx), _ptr(vals), _ptr(w), _ptr(rnorm), _ptr(dz),
          Np * D, _ptr(dvals), _ptr(dh, off), st)
    assert rel_err(dvals, dvals_ref) < 1e-4
    assert rel_err(dh, dh_ref.float()) < 1e-2
    assert float(dh[:, 0].abs().max()) == 0.0
