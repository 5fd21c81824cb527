"""Host-side operators over the libgvit C ABI (``include/gvit.h``).

Each function here is one stage of the graph-augmented ViT hot path:

* ``attention_core``   - a2, /root/reference/src/models/vit.py:59-69
* ``layer_norm``       - a5, nn.LayerNorm at vit.py:103,108,154
* ``dropout_add``      - a5, proj_drop + residual at vit.py:71,117
* ``gelu_dropout``     - Mlp activation edge, nn.GELU + nn.Dropout at vit.py:84,92
* ``knn_graph``        - a7, SURVEY.md section 9 G1-G3 (no reference symbol)
* ``patch_graph``      - a7+a8 as one differentiable op, section 9 G0-G6 (no reference symbol)
* ``linear`` / ``linear_gelu_dropout`` / ``linear_dropout_add`` / ``mlp_fused`` - f1, the nn.Linear layers of the block
  (vit.py:50,52,83-94) with their edges: fused tcgen05 GEMMs for fc1, the attention projection and fc2's input gradient,
  library GEMMs elsewhere; parameter shadows (one multi-tensor cast per step) and fp32 weight gradients
* ``patch_embed_tokens`` - f4, PatchEmbed + CLS + pos_embed + pos_drop (vit.py:25-36, 207-212)

PyTorch is plumbing only: it owns the device memory and the stream, the arithmetic runs in the CUDA
library.  There is no fallback: a CPU tensor, an unsupported dtype or a missing ``libgvit.so`` raises.
Under ``torch.autocast`` the ops run in bf16 (fp16 autocast, which the reference trainer uses at
/root/reference/src/training/trainer.py:101, is also executed in bf16: same 16-bit storage, no loss
scaling hazards).
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import GVIT_BF16, GVIT_COLSUM_CHUNKS, GVIT_F32, GVIT_LN_PARTIALS

__all__ = ["attention_core", "layer_norm", "pre_norm", "linear", "colsum", "linear_dropout_add", "linear_gelu_dropout", "dropout_add", "gelu_dropout", "knn_graph", "graph_reverse", "patch_graph",
           "agg_gather", "mlp_fused", "patch_embed_tokens", "refresh_shadows", "invalidate_shadows", "set_rng_offset_tensor", "launch_count", "reset_launch_count"]

# kernels launched through the C ABI since the last reset (bench.py reports it as gpu_launches)
_LAUNCHES = {"n": 0}
_KERNELS_PER_CALL = {
    "gvit_knn_fwd": 1, "gvit_graph_reverse": 1, "gvit_knn_bwd": 1, "gvit_agg_gather_fwd": 1, "gvit_agg_fwd": 1,
    "gvit_agg_bwd": 2, "gvit_graph_bwd": 2, "gvit_attn_fwd": 1, "gvit_attn_bwd": 1, "gvit_layernorm_fwd": 1, "gvit_layernorm_bwd": 2,
    "gvit_colsum": 2, "gvit_dropout_residual_fwd": 1, "gvit_dropout_bwd": 1, "gvit_gelu_dropout_fwd": 1, "gvit_gelu_dropout_bwd": 1,
    "gvit_patchify": 1, "gvit_embed_assemble": 1, "gvit_linear_gelu_dropout_fwd": 1, "gvit_linear_dropout_residual_fwd": 1, "gvit_linear_gelu_dropout_bwd": 2,
}


def launch_count() -> int:
    return _LAUNCHES["n"]


def reset_launch_count() -> None:
    _LAUNCHES["n"] = 0


def _call(name, *args):
    _lib.call(name, *args)
    _LAUNCHES["n"] += _KERNELS_PER_CALL.get(name, 1)


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return GVIT_F32
    if t.dtype == torch.bfloat16:
        return GVIT_BF16
    raise TypeError(f"libgvit computes in float32 or bfloat16, got {t.dtype}")


def _check_cuda(*tensors):
    """Every tensor argument on ONE CUDA device, and that device current: the C side launches on the current device's
    stream and builds its TMA descriptors there (one process per GPU) - a model on cuda:1 under a current
    device cuda:0 would otherwise launch against foreign pointers with no stream ordering."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("libgvit operators run on CUDA tensors only (there is no CPU fallback); "
                               f"got a tensor on {t.device}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"libgvit operators need all tensors on one device; got {dev} and {t.device}")
    if dev is not None and dev.index != torch.cuda.current_device():
        raise RuntimeError(f"libgvit operators run on the CURRENT CUDA device (cuda:{torch.cuda.current_device()}); the tensors "
                           f"are on {dev} - wrap the call in `with torch.cuda.device({dev.index}):` or torch.cuda.set_device")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t, byte_offset=0):
    return None if t is None else t.data_ptr() + byte_offset


def _autocast_dtype(t: torch.Tensor):
    """bf16 when CUDA autocast is on (any 16-bit autocast dtype is executed as bf16), else the tensor's dtype."""
    if torch.is_autocast_enabled("cuda"):
        return torch.bfloat16
    if t.dtype not in (torch.float32, torch.bfloat16):
        if t.dtype == torch.float16:
            return torch.bfloat16
        raise TypeError(f"unsupported dtype {t.dtype}")
    return t.dtype


# ------------------------------------------------------------------------------------------------
# Parameter shadows: the 16-bit copies of the fp32 master parameters that autocast would make one `.to()` at a time
# (335 cast kernels per forward of ViT-B + graph) are made by ONE multi-tensor copy per step, and the operators below
# take the MASTER parameter as their autograd input: their backward returns fp32 gradients directly (the weight-gradient
# GEMM accumulates and writes fp32), so there is no 16-bit gradient and no cast on the way back either.
# ------------------------------------------------------------------------------------------------
_SHADOWS: dict = {}          # id(param) -> (weakref(param), version, data_ptr, shadow tensor, epoch)
_SHADOW_EPOCH = {"n": 0}


def invalidate_shadows() -> None:
    """Declare every shadow stale.  Needed when parameters change without their `_version` moving: a CUDA-graph replay
    of an optimizer step updates the masters on the device without executing any Python."""
    _SHADOW_EPOCH["n"] += 1



def refresh_shadows(params, dtype: torch.dtype, force: bool = False) -> None:
    """Bring the `dtype` shadows of `params` up to date with one multi-tensor copy.

    force=False copies stale or missing shadows only, where "stale" is judged by `_version`, `data_ptr` and the epoch of
    `invalidate_shadows` - which in-place writes through `.data` (EMA, clamping, `dist.broadcast(p.data)`) do NOT move.
    force=True re-copies every shadow (same buffers): the model's forward uses it, so a shadow can never outlive a
    parameter update however it was made (0.17 GB read + 0.17 GB written for ViT-B: ~60 us per forward)."""
    import weakref
    src, dst = [], []
    for prm in params:
        if prm is None or prm.dtype == dtype or not prm.is_cuda:
            continue
        e = _SHADOWS.get(id(prm))
        if e is not None and e[0]() is prm and e[3].dtype == dtype and e[3].shape == prm.shape:
            if not force and e[1] == prm._version and e[2] == prm.data_ptr() and e[4] == _SHADOW_EPOCH["n"]:
                continue
            sh = e[3]
        else:
            sh = torch.empty_like(prm, dtype=dtype, memory_format=torch.contiguous_format)
        _SHADOWS[id(prm)] = (weakref.ref(prm, lambda _r, k=id(prm): _SHADOWS.pop(k, None)), prm._version, prm.data_ptr(), sh,
                             _SHADOW_EPOCH["n"])
        src.append(prm.detach())
        dst.append(sh)
    if src:
        with torch.no_grad():
            torch._foreach_copy_(dst, src)


def _shadow(prm, dtype: torch.dtype):
    """`prm` in the compute dtype, without an autograd edge: the parameter itself, its fresh shadow, or a one-off cast."""
    if prm is None:
        return None
    t = prm.detach()
    if t.dtype == dtype:
        return t if t.is_contiguous() else t.contiguous()
    e = _SHADOWS.get(id(prm))
    if (e is not None and e[0]() is prm and e[1] == prm._version and e[2] == prm.data_ptr() and e[3].dtype == dtype
            and e[4] == _SHADOW_EPOCH["n"]):
        return e[3]
    return t.to(dtype).contiguous()


# ------------------------------------------------------------------------------------------------
# a1 / f1: the dense GEMMs of every nn.Linear - forward, input gradient, weight gradient - through gvit_linear_gemm (one
# persistent 2-SM tcgen05 kernel, csrc/gemm2_tc.cu) whenever the operands are bf16 with an output width that is a multiple
# of 256; the exact-fp32 parity path and the 14-logit head (vit.py:176) stay torch matmuls.
# ------------------------------------------------------------------------------------------------
_GEMM2 = {"on": os.environ.get("GVIT_GEMM2", "1") != "0"}             # GVIT_GEMM2=0: library GEMMs (A/B switch)


def _gemm2_operand(t: torch.Tensor) -> bool:
    return (t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 8 == 0 and t.stride(0) >= t.shape[1]
            and t.data_ptr() % 16 == 0)


def _gemm2(a, a_t, b, b_t, M, N, K, bias, out):
    ws_bytes = 0
    ws = None
    if out.dtype == torch.float32:
        ws_bytes = int(_lib.load().gvit_linear_gemm_ws_bytes(M, N, K))
        if ws_bytes:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=out.device)
    _call("gvit_linear_gemm", _ptr(a), a_t, a.stride(0), _ptr(b), b_t, b.stride(0), M, N, K, _ptr(bias),
          GVIT_F32 if out.dtype == torch.float32 else GVIT_BF16, _ptr(out), out.stride(0), _ptr(ws), ws_bytes, _stream())
    return out


def _mm_nt(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """``F.linear(x, w, bias)``: y = x w^T + bias over the last dimension of x (w is (N, K))."""
    N, K = w.shape
    if (_GEMM2["on"] and x.is_cuda and x.dtype == torch.bfloat16 and N % 256 == 0 and K % 64 == 0 and x.is_contiguous()
            and _gemm2_operand(w) and (bias is None or (bias.dtype == torch.bfloat16 and bias.is_contiguous() and bias.data_ptr() % 16 == 0))):
        x2 = x.view(-1, K)
        if _gemm2_operand(x2):
            y = torch.empty(x.shape[:-1] + (N,), dtype=torch.bfloat16, device=x.device)
            _gemm2(x2, 0, w, 0, x2.shape[0], N, K, bias, y.view(-1, N))
            return y
    return F.linear(x, w, bias)


def _mm_nn(dy2: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``dy2 @ w``: the input gradient of a Linear, w (N, K) read as stored (an MN-major B operand, no transpose copy)."""
    N, K = w.shape
    if _GEMM2["on"] and dy2.is_cuda and K % 256 == 0 and N % 64 == 0 and _gemm2_operand(dy2) and _gemm2_operand(w):
        dx = torch.empty((dy2.shape[0], K), dtype=torch.bfloat16, device=dy2.device)
        return _gemm2(dy2, 0, w, 1, dy2.shape[0], K, N, None, dx)
    return dy2 @ w


def _wgrad(dy2: torch.Tensor, x2: torch.Tensor, master_dtype: torch.dtype) -> torch.Tensor:
    """dW = dy2^T x2 in the MASTER parameter's dtype: an fp32 master over 16-bit operands gets the GEMM's fp32
    accumulator written out as is (no bf16 rounding of the gradient, no cast kernel afterwards).  Both operands are read
    as stored (MN-major); the reduction over all rows is split into pieces added in a fixed order (deterministic)."""
    if master_dtype == torch.float32 and dy2.dtype != torch.float32:
        N, K = dy2.shape[1], x2.shape[1]
        if _GEMM2["on"] and dy2.is_cuda and K % 256 == 0 and N % 64 == 0 and _gemm2_operand(dy2) and _gemm2_operand(x2):
            dw = torch.empty((N, K), dtype=torch.float32, device=dy2.device)
            return _gemm2(dy2, 1, x2, 1, N, K, dy2.shape[0], None, dw)
        return torch.mm(dy2.t(), x2, out_dtype=torch.float32)
    return (dy2.t() @ x2).to(master_dtype)


def _colsum_ws(rows: int, D: int, device) -> torch.Tensor:
    """Workspace of gvit_colsum / the *_bwd column sums: an upper bound of the row chunks the launchers use (at most 12
    per SM and column block, never more than GVIT_COLSUM_CHUNKS or rows / 32) x D floats."""
    cb = (D + 255) // 256
    n = min(GVIT_COLSUM_CHUNKS, (12 * _sm_count(device) + cb - 1) // cb + 1, max(1, (rows + 31) // 32) + 1)
    return torch.empty(n * D, dtype=torch.float32, device=device)


_SMS: dict = {}


def _sm_count(device) -> int:
    i = torch.device(device).index
    i = torch.cuda.current_device() if i is None else i
    if i not in _SMS:
        _SMS[i] = torch.cuda.get_device_properties(i).multi_processor_count
    return _SMS[i]


# ------------------------------------------------------------------------------------------------
# a2: attention core
# ------------------------------------------------------------------------------------------------
class _AttentionCore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, num_heads, scale):
        B, N, C3 = qkv.shape
        dh = C3 // (3 * num_heads)
        out = torch.empty((B, N, num_heads * dh), dtype=qkv.dtype, device=qkv.device)
        lse = torch.empty((B, num_heads, N), dtype=torch.float32, device=qkv.device)
        _call("gvit_attn_fwd", _ptr(qkv), B, N, num_heads, dh, float(scale), _dtype_code(qkv), _ptr(out), _ptr(lse),
              _stream())
        ctx.save_for_backward(qkv, out, lse)
        ctx.meta = (num_heads, dh, float(scale))
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        H, dh, scale = ctx.meta
        B, N, _ = qkv.shape
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        delta = torch.empty_like(lse)
        _call("gvit_attn_bwd", _ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), B, N, H, dh, scale, _dtype_code(qkv),
              _ptr(delta), _ptr(dqkv), _stream())
        return dqkv, None, None


def attention_core(qkv: torch.Tensor, num_heads: int, scale: float) -> torch.Tensor:
    """softmax((q k^T) * scale) v on the packed projection output of vit.py:59.

    qkv: (B, N, 3*H*dh) laid out as (B, N, 3, H, dh); returns (B, N, H*dh), head-major - the tensor
    vit.py:69 produces after its transpose + reshape.  The (B,H,N,N) score tensor is never materialised.
    """
    _check_cuda(qkv)
    if qkv.dim() != 3 or qkv.shape[-1] % (3 * num_heads):
        raise ValueError(f"qkv must be (B, N, 3*H*dh); got {tuple(qkv.shape)} with H={num_heads}")
    dt = _autocast_dtype(qkv)
    with torch.autocast("cuda", enabled=False):
        return _AttentionCore.apply(qkv.to(dt).contiguous(), int(num_heads), float(scale))


# ------------------------------------------------------------------------------------------------
# a1 / a3: the Linear layers around the fused kernels (library GEMMs; libgvit supplies the bias gradient)
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def colsum(x2: torch.Tensor, skip_period: int = 0) -> torch.Tensor:
    """fp32 column sums of a contiguous (rows, D) tensor - deterministic two-pass reduction (gvit_colsum).
    ``skip_period`` > 0 leaves rows r with r % skip_period == 0 out (the CLS rows of a (B, 1+Np, D) tensor)."""
    _check_cuda(x2)
    rows, D = x2.shape
    out = torch.empty(D, dtype=torch.float32, device=x2.device)
    ws = _colsum_ws(rows, D, x2.device)
    _call("gvit_colsum", _ptr(x2), rows, D, _dtype_code(x2), int(skip_period), _ptr(out), _ptr(ws), _stream())
    return out


class _Linear(torch.autograd.Function):
    """y = x W^T + b (vit.py:59,70,90,93).  The three GEMMs stay library GEMMs (cuBLASLt); the bias
    gradient is a libgvit column sum instead of at::sum (3.7 ms of a 53 ms training step)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        w = _shadow(weight, x.dtype)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.bias_dtype = bias.dtype if bias is not None else None
        ctx.w_dtype = weight.dtype
        return _mm_nt(x, w, _shadow(bias, x.dtype))

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = _mm_nn(dy2, weight).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = _wgrad(dy2, x.reshape(-1, x.shape[-1]), ctx.w_dtype) if ctx.needs_input_grad[1] else None
        db = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = (colsum(dy2) if dy2.shape[1] % 8 == 0 else dy2.float().sum(0)).to(ctx.bias_dtype)
        return dx, dw, db


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None) -> torch.Tensor:
    """``F.linear`` with autocast semantics (16-bit autocast runs in bf16) and a libgvit bias gradient."""
    _check_cuda(x, weight, bias)
    dt = _autocast_dtype(x)
    with torch.autocast("cuda", enabled=False):
        return _Linear.apply(x.to(dt), weight, bias)      # the master parameters themselves: their gradients come back in their dtype


# ------------------------------------------------------------------------------------------------
# a5: LayerNorm, dropout + residual, GELU + dropout
# ------------------------------------------------------------------------------------------------
_TORCH_DT = {GVIT_F32: torch.float32, GVIT_BF16: torch.bfloat16}


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, y_code):
        D = x.shape[-1]
        rows = x.numel() // D
        y = torch.empty(x.shape, dtype=_TORCH_DT[y_code], device=x.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        ctx.w_dtype = weight.dtype
        weight, bias = _shadow(weight, x.dtype), _shadow(bias, x.dtype)
        _call("gvit_layernorm_fwd", _ptr(x), _ptr(weight), _ptr(bias), rows, D, float(eps), _dtype_code(x), y_code,
              _ptr(y), _ptr(mean), _ptr(rstd), _stream())
        ctx.save_for_backward(x, weight, mean, rstd)
        ctx.y_code = y_code
        return y

    @staticmethod
    def backward(ctx, dy):
        dx, dgamma, dbeta = _layer_norm_backward(ctx, dy, None)
        return dx, dgamma, dbeta, None, None


def _layer_norm_backward(ctx, dy, dx_add):
    x, weight, mean, rstd = ctx.saved_tensors
    D = x.shape[-1]
    rows = x.numel() // D
    dy = dy.contiguous()
    dx = torch.empty_like(x)
    if dx_add is not None:
        dx_add = dx_add.to(x.dtype).contiguous()
    # two separate tensors: AccumulateGrad steals a whole tensor but has to clone a view
    dgamma = torch.empty(D, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(D, dtype=torch.float32, device=x.device)
    ws = torch.empty(2 * GVIT_LN_PARTIALS * D, dtype=torch.float32, device=x.device)
    _call("gvit_layernorm_bwd", _ptr(dy), _ptr(x), _ptr(weight), _ptr(mean), _ptr(rstd), rows, D, _dtype_code(x),
          ctx.y_code, _ptr(dx_add), _ptr(dx), _ptr(dgamma), _ptr(dbeta), _ptr(ws), _stream())
    return dx, dgamma.to(ctx.w_dtype), dbeta.to(ctx.w_dtype)


class _PreNorm(torch.autograd.Function):
    """(x, LayerNorm(x)): the first output is x itself, handed to the residual add of the sub-layer.  Because both
    uses of x leave through ONE node, its backward receives the residual-path gradient together with dy and the
    LayerNorm backward kernel adds it to dx in the same pass (autograd would otherwise launch a separate add)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, y_code):
        y = _LayerNorm.forward(ctx, x, weight, bias, eps, y_code)
        return x.view_as(x), y

    @staticmethod
    def backward(ctx, dres, dy):
        if dy is None:                      # the branch was unused: only the residual path carries gradient
            return dres, None, None, None, None
        dx, dgamma, dbeta = _layer_norm_backward(ctx, dy, dres)
        return dx, dgamma, dbeta, None, None


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.LayerNorm over the last dimension (affine), statistics in fp32.

    Outside autocast the output has x's dtype.  Under autocast an fp32 x (fp32 residual stream) yields a bf16
    output - the cast the consuming Linear / graph / attention op would do anyway, folded into the kernel.
    """
    _check_cuda(x, weight, bias)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.to(torch.bfloat16 if x.dtype == torch.float16 else torch.float32)
    y_code = GVIT_BF16 if (x.dtype == torch.bfloat16 or torch.is_autocast_enabled("cuda")) else GVIT_F32
    with torch.autocast("cuda", enabled=False):
        return _LayerNorm.apply(x.contiguous(), weight, bias, float(eps), y_code)


def pre_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5):
    """``(x, layer_norm(x))`` for a pre-norm residual sub-layer ``x + f(LN(x))`` (vit.py:117-118): use the returned x
    for the residual add.  Same values as ``layer_norm``; the backward folds the residual-path gradient into the
    LayerNorm backward kernel."""
    _check_cuda(x, weight, bias)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.to(torch.bfloat16 if x.dtype == torch.float16 else torch.float32)
    y_code = GVIT_BF16 if (x.dtype == torch.bfloat16 or torch.is_autocast_enabled("cuda")) else GVIT_F32
    with torch.autocast("cuda", enabled=False):
        return _PreNorm.apply(x.contiguous(), weight, bias, float(eps), y_code)


# Device-resident Philox offset for CUDA-graph capture: a captured launch bakes its host arguments (seed, offset)
# in, so graph replays would repeat the dropout masks.  While an offset tensor is installed every dropout kernel adds
# the uint64 it holds to its counter; the captured step advances it on the device (see step.CapturedTrainStep).
_RNG_OFFSET = {"t": None}


def set_rng_offset_tensor(t: torch.Tensor | None) -> None:
    """Install (or remove, with None) a 1-element int64 CUDA tensor whose value is added to every dropout counter."""
    if t is not None and (t.dtype != torch.int64 or t.numel() != 1 or not t.is_cuda):
        raise ValueError("the rng offset must be a 1-element int64 CUDA tensor")
    _RNG_OFFSET["t"] = t


def _rng_offset_ptr():
    t = _RNG_OFFSET["t"]
    return None if t is None else t.data_ptr()


def _draw_seed() -> int:
    # CPU generator: follows torch.manual_seed, costs no device synchronisation
    return int(torch.randint(0, 2 ** 62, (1,)).item())


class _DropoutAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, resid, p, seed):
        n = y.numel()
        out = torch.empty_like(y if resid is None else resid)
        mask = torch.empty(n // 8, dtype=torch.uint8, device=y.device) if p > 0 else None
        _call("gvit_dropout_residual_fwd", _ptr(y), _ptr(resid), n, float(p), int(seed), 0, _rng_offset_ptr(), _dtype_code(out),
              _dtype_code(y), _ptr(out), _ptr(mask), _stream())
        ctx.p = p
        ctx.has_resid = resid is not None
        ctx.y_dtype = y.dtype
        if p > 0:
            ctx.save_for_backward(mask)
        return out

    @staticmethod
    def backward(ctx, dout):
        dresid = dout if ctx.has_resid else None
        if ctx.p == 0:
            return dout.to(ctx.y_dtype), dresid, None, None
        (mask,) = ctx.saved_tensors
        dout = dout.contiguous()
        dy = torch.empty(dout.shape, dtype=ctx.y_dtype, device=dout.device)
        _call("gvit_dropout_bwd", _ptr(dout), _ptr(mask), dout.numel(), float(ctx.p), _dtype_code(dout),
              _dtype_code(dy), _ptr(dy), 0, 0, None, None, _stream())
        return dy, dresid, None, None


def dropout_add(y: torch.Tensor, resid: torch.Tensor | None, p: float, training: bool) -> torch.Tensor:
    """``resid + dropout(y, p)`` in one pass (Philox-4x32-7 keep mask, stored as one bit per element).

    ``resid`` may be None (plain dropout).  A bf16 branch ``y`` may be added onto an fp32 stream ``resid``.
    The seed of each call is drawn from PyTorch's CPU generator, so ``torch.manual_seed`` makes runs
    reproducible without a device synchronisation.
    """
    _check_cuda(y, resid)
    p = float(p) if training else 0.0
    if p == 0.0 and resid is None:
        return y
    if y.dtype not in (torch.float32, torch.bfloat16):
        y = y.to(torch.bfloat16)
    if resid is not None:
        if resid.dtype not in (torch.float32, torch.bfloat16):
            resid = resid.to(torch.bfloat16)
        if resid.dtype == torch.bfloat16 and y.dtype == torch.float32:
            resid = resid.float()                    # only (stream fp32, branch bf16) is a mixed pairing
    if y.numel() % 8:
        raise ValueError("dropout_add needs a multiple of 8 elements")
    seed = _draw_seed() if p > 0 else 0
    with torch.autocast("cuda", enabled=False):
        return _DropoutAdd.apply(y.contiguous(), None if resid is None else resid.contiguous(), p, seed)


class _GeluDropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, p, seed):
        n = u.numel()
        out = torch.empty_like(u)
        mask = torch.empty(n // 8, dtype=torch.uint8, device=u.device) if p > 0 else None
        _call("gvit_gelu_dropout_fwd", _ptr(u), n, float(p), int(seed), 0, _rng_offset_ptr(), _dtype_code(u), _ptr(out), _ptr(mask),
              _stream())
        ctx.p = p
        ctx.save_for_backward(u, mask)
        return out

    @staticmethod
    def backward(ctx, dout):
        u, mask = ctx.saved_tensors
        dout = dout.contiguous()
        du = torch.empty_like(u)
        _call("gvit_gelu_dropout_bwd", _ptr(dout), _ptr(u), _ptr(mask), u.numel(), float(ctx.p), _dtype_code(u),
              _ptr(du), 0, None, None, _stream())
        return du, None, None


def gelu_dropout(u: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    """``dropout(gelu(u), p)`` with the exact-erf GELU of nn.GELU (vit.py:84,92) in one pass; the backward
    recomputes GELU' from u instead of keeping the activation."""
    _check_cuda(u)
    p = float(p) if training else 0.0
    dt = _autocast_dtype(u)
    if u.numel() % 8:
        raise ValueError("gelu_dropout needs a multiple of 8 elements")
    seed = _draw_seed() if p > 0 else 0
    with torch.autocast("cuda", enabled=False):
        return _GeluDropout.apply(u.to(dt).contiguous(), p, seed)


# ------------------------------------------------------------------------------------------------
# Linear + edge, as one autograd node each: the edge's backward kernel also yields the Linear's bias gradient
# ------------------------------------------------------------------------------------------------
def _linear_grads(ctx_needs, x, weight, dy2, w_dtype):
    dx = _mm_nn(dy2, weight).view(x.shape) if ctx_needs[0] else None
    dw = _wgrad(dy2, x.reshape(-1, x.shape[-1]), w_dtype) if ctx_needs[1] else None
    return dx, dw


class _LinearDropoutAdd(torch.autograd.Function):
    """out = resid + dropout(x W^T + b, p)  (proj + proj_drop + residual, vit.py:70-71,117; fc2 + drop + residual,
    vit.py:93-94,118).  Backward: ONE pass over dout applies the keep mask and accumulates the column sums that are
    the bias gradient; the two GEMM gradients are library GEMMs."""

    @staticmethod
    def forward(ctx, x, weight, bias, resid, p, seed):
        ctx.bias_dtype = bias.dtype if bias is not None else None
        ctx.w_dtype = weight.dtype
        weight = _shadow(weight, x.dtype)
        N, K = weight.shape
        if (x.dtype == torch.bfloat16 and resid is not None and K <= _FUSED_RESID_MAX_K
                and fused_fc1_available(N, K) and x.is_contiguous()):
            # proj + proj_drop + residual add as ONE tcgen05 GEMM;
            # resid / out are bf16, or fp32 when the residual stream is kept in fp32 (autocast semantics)
            M = x.numel() // K
            out = torch.empty_like(resid)
            mask = torch.empty(M * N // 8, dtype=torch.uint8, device=x.device) if p > 0 else None
            _call("gvit_linear_dropout_residual_fwd", _ptr(x), _ptr(weight), _ptr(_shadow(bias, x.dtype)), _ptr(resid), M, N, K,
                  float(p), int(seed), 0, _rng_offset_ptr(), GVIT_BF16, _dtype_code(resid), _ptr(out), _ptr(mask), _stream())
            y_dtype = x.dtype
        else:
            y = _mm_nt(x, weight, _shadow(bias, x.dtype))
            n = y.numel()
            out = torch.empty_like(y if resid is None else resid)
            mask = torch.empty(n // 8, dtype=torch.uint8, device=y.device) if p > 0 else None
            _call("gvit_dropout_residual_fwd", _ptr(y), _ptr(resid), n, float(p), int(seed), 0, _rng_offset_ptr(), _dtype_code(out),
                  _dtype_code(y), _ptr(out), _ptr(mask), _stream())
            y_dtype = y.dtype
        ctx.save_for_backward(x, weight, mask)
        ctx.p, ctx.has_bias, ctx.has_resid, ctx.y_dtype = p, bias is not None, resid is not None, y_dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, mask = ctx.saved_tensors
        dout = dout.contiguous()
        Dn = dout.shape[-1]
        dy = torch.empty(dout.shape, dtype=ctx.y_dtype, device=dout.device)
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        db = torch.empty(Dn, dtype=torch.float32, device=dout.device) if want_db else None
        ws = _colsum_ws(dout.numel() // Dn, Dn, dout.device) if want_db else None
        if ctx.p > 0 or want_db:
            _call("gvit_dropout_bwd", _ptr(dout), _ptr(mask), dout.numel(), float(ctx.p), _dtype_code(dout),
                  _dtype_code(dy), _ptr(dy), Dn, 0, _ptr(db), _ptr(ws), _stream())
        else:
            dy = dout.to(ctx.y_dtype)
        dx, dw = _linear_grads(ctx.needs_input_grad, x, weight, dy.view(-1, Dn), ctx.w_dtype)
        return dx, dw, (db.to(ctx.bias_dtype) if want_db else None), (dout if ctx.has_resid else None), None, None


class _LinearGeluDropout(torch.autograd.Function):
    """out = dropout(gelu(x W^T + b), p)  (fc1 + GELU + drop, vit.py:90-92); the pre-activation is saved, GELU' is
    recomputed, and the backward pass over dout also accumulates fc1's bias gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias, p, seed):
        ctx.bias_dtype = bias.dtype if bias is not None else None
        ctx.w_dtype = weight.dtype
        weight = _shadow(weight, x.dtype)
        N, K = weight.shape
        if x.dtype == torch.bfloat16 and fused_fc1_available(N, K) and x.is_contiguous():
            # one persistent tcgen05 GEMM: bias + GELU + dropout + both stores run in its epilogue warps
            M = x.numel() // K
            u = torch.empty(x.shape[:-1] + (N,), dtype=x.dtype, device=x.device)
            out = torch.empty_like(u)
            mask = torch.empty(M * N // 8, dtype=torch.uint8, device=x.device) if p > 0 else None
            _call("gvit_linear_gelu_dropout_fwd", _ptr(x), _ptr(weight), _ptr(_shadow(bias, x.dtype)), M, N, K, float(p),
                  int(seed), 0, _rng_offset_ptr(), GVIT_BF16, 0, _ptr(u), _ptr(out), _ptr(mask), _stream())
        else:
            u = _mm_nt(x, weight, _shadow(bias, x.dtype))
            n = u.numel()
            out = torch.empty_like(u)
            mask = torch.empty(n // 8, dtype=torch.uint8, device=u.device) if p > 0 else None
            _call("gvit_gelu_dropout_fwd", _ptr(u), n, float(p), int(seed), 0, _rng_offset_ptr(), _dtype_code(u), _ptr(out), _ptr(mask), _stream())
        ctx.save_for_backward(x, weight, u, mask)
        ctx.p, ctx.has_bias = p, bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, u, mask = ctx.saved_tensors
        dout = dout.contiguous()
        Dn = u.shape[-1]
        du = torch.empty_like(u)
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        db = torch.empty(Dn, dtype=torch.float32, device=u.device) if want_db else None
        ws = _colsum_ws(u.numel() // Dn, Dn, u.device) if want_db else None
        _call("gvit_gelu_dropout_bwd", _ptr(dout), _ptr(u), _ptr(mask), u.numel(), float(ctx.p), _dtype_code(u),
              _ptr(du), Dn, _ptr(db), _ptr(ws), _stream())
        dx, dw = _linear_grads(ctx.needs_input_grad, x, weight, du.view(-1, Dn), ctx.w_dtype)
        return dx, dw, (db.to(ctx.bias_dtype) if want_db else None), None, None


class _MlpFused(torch.autograd.Function):
    """out = resid + dropout(fc2(dropout(gelu(fc1(x)))))  - the whole Mlp branch of vit.py:88-94,118 as one autograd node, so
    that the backward can run fc2's input-gradient GEMM fused with the GELU / dropout backward and fc1's bias gradient
    (gvit_linear_gelu_dropout_bwd): the (M, 4D) gradient of the hidden activation never makes an HBM round trip.
    Forward: fused fc1 GEMM (gvit_linear_gelu_dropout_fwd, saving the backward factor), fused fc2 + drop + residual GEMM."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, resid, p, seed1, seed2):
        ctx.dtypes = (w1.dtype, None if b1 is None else b1.dtype, w2.dtype, None if b2 is None else b2.dtype)
        w1s, w2s = _shadow(w1, x.dtype), _shadow(w2, x.dtype)
        Nh, K = w1s.shape
        M = x.numel() // K
        st = _stream()
        # what the backward needs of fc1 + GELU + drop is the FACTOR keep * gelu'(u) / (1 - p): the fused kernel writes it in
        # place of the pre-activation (no keep mask), so that the backward GEMM's epilogue is one multiply; nothing at all
        # is saved when no gradient is wanted (inference: one (M, 4D) store less per layer)
        want_grad = any(ctx.needs_input_grad[:5])
        mode = 1 if _MLP_FACTOR["on"] else 0
        h = torch.empty(x.shape[:-1] + (Nh,), dtype=x.dtype, device=x.device)
        u = torch.empty_like(h) if want_grad else None
        mask1 = torch.empty(M * Nh // 8, dtype=torch.uint8, device=x.device) if (p > 0 and want_grad and mode == 0) else None
        _call("gvit_linear_gelu_dropout_fwd", _ptr(x), _ptr(w1s), _ptr(_shadow(b1, x.dtype)), M, Nh, K, float(p), int(seed1), 0,
              _rng_offset_ptr(), GVIT_BF16, mode, _ptr(u), _ptr(h), _ptr(mask1), st)
        ctx.mode = mode
        D2 = w2s.shape[0]
        if resid is not None and Nh <= _FUSED_RESID_MAX_K and fused_fc1_available(D2, Nh):
            # fc2 + drop + residual add (vit.py:93-94,118) as ONE tcgen05 GEMM: bias, keep mask and the add run in its epilogue
            out = torch.empty_like(resid)
            mask2 = torch.empty(M * D2 // 8, dtype=torch.uint8, device=x.device) if p > 0 else None
            _call("gvit_linear_dropout_residual_fwd", _ptr(h), _ptr(w2s), _ptr(_shadow(b2, x.dtype)), _ptr(resid), M, D2, Nh, float(p),
                  int(seed2), 0, _rng_offset_ptr(), GVIT_BF16, _dtype_code(resid), _ptr(out), _ptr(mask2), st)
            y_dtype = x.dtype
        else:
            y = _mm_nt(h, w2s, _shadow(b2, x.dtype))
            n = y.numel()
            out = torch.empty_like(y if resid is None else resid)
            mask2 = torch.empty(n // 8, dtype=torch.uint8, device=x.device) if p > 0 else None
            _call("gvit_dropout_residual_fwd", _ptr(y), _ptr(resid), n, float(p), int(seed2), 0, _rng_offset_ptr(), _dtype_code(out),
                  _dtype_code(y), _ptr(out), _ptr(mask2), st)
            y_dtype = y.dtype
        ctx.save_for_backward(x, w1s, w2s, u, h, mask1, mask2)
        ctx.p, ctx.has_resid, ctx.y_dtype = p, resid is not None, y_dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w1s, w2s, u, h, mask1, mask2 = ctx.saved_tensors
        w1_dt, b1_dt, w2_dt, b2_dt = ctx.dtypes
        Nh, K = w1s.shape
        D2 = w2s.shape[0]
        M = x.numel() // K
        st = _stream()
        dout = dout.contiguous()
        # fc2: dropout backward + bias gradient in one pass, weight gradient as a library GEMM
        dy = torch.empty(dout.shape, dtype=ctx.y_dtype, device=dout.device)
        db2 = torch.empty(D2, dtype=torch.float32, device=dout.device) if b2_dt is not None else None
        ws2 = _colsum_ws(M, D2, dout.device) if db2 is not None else None
        if ctx.p > 0 or db2 is not None:
            _call("gvit_dropout_bwd", _ptr(dout), _ptr(mask2), dout.numel(), float(ctx.p), _dtype_code(dout), _dtype_code(dy),
                  _ptr(dy), D2, 0, _ptr(db2), _ptr(ws2), st)
        else:
            dy = dout.to(ctx.y_dtype)
        dy2 = dy.view(M, D2)
        dw2 = _wgrad(dy2, h.view(M, Nh), w2_dt) if ctx.needs_input_grad[3] else None
        # fc2 input gradient + GELU' + keep mask + fc1 bias gradient: ONE tcgen05 GEMM, W2 read as stored (MN-major B)
        du = torch.empty_like(u)
        db1 = torch.empty(Nh, dtype=torch.float32, device=dout.device)
        rows = _lib.load().gvit_linear_gelu_dropout_bwd_ws_rows(M)
        part = torch.empty(rows * Nh, dtype=torch.float32, device=dout.device)
        _call("gvit_linear_gelu_dropout_bwd", _ptr(dy2), _ptr(w2s), _ptr(u), _ptr(mask1), M, Nh, D2, float(ctx.p), GVIT_BF16, ctx.mode,
              _ptr(du), _ptr(db1), _ptr(part), st)
        du2 = du.view(M, Nh)
        dx = _mm_nn(du2, w1s).view(x.shape) if ctx.needs_input_grad[0] else None
        dw1 = _wgrad(du2, x.reshape(M, K), w1_dt) if ctx.needs_input_grad[1] else None
        return (dx, dw1, (db1.to(b1_dt) if b1_dt is not None and ctx.needs_input_grad[2] else None), dw2,
                (db2.to(b2_dt) if db2 is not None and ctx.needs_input_grad[4] else None),
                (dout if ctx.has_resid else None), None, None, None)


_MLP_FACTOR = {"on": os.environ.get("GVIT_MLP_FACTOR", "1") != "0"}   # GVIT_MLP_FACTOR=0: save the pre-activation + keep mask (A/B switch)
_MLP_FUSED = {"on": os.environ.get("GVIT_MLP_FUSED", "1") != "0"}      # GVIT_MLP_FUSED=0: two-op composition (A/B switch)


def mlp_fused_available(x, w1, w2, resid) -> bool:
    """True when the whole-Mlp node applies: bf16 compute, hidden width a multiple of 256, widths multiples of 64 (the
    residual stream may be bf16 or fp32)."""
    dt = _autocast_dtype(x)
    return (_MLP_FUSED["on"] and dt == torch.bfloat16 and x.is_cuda and w1.dim() == 2 and w2.dim() == 2
            and w2.shape[1] == w1.shape[0] and w2.shape[0] % 8 == 0
            and (resid is None or resid.dtype in (torch.bfloat16, torch.float32))
            and fused_fc1_available(w1.shape[0], w1.shape[1]) and fused_fc1_available(w1.shape[0], w2.shape[0]))


def mlp_fused(x, w1, b1, w2, b2, resid, p: float, training: bool):
    """``resid + dropout(linear(dropout(gelu(linear(x, w1, b1))), w2, b2))`` as one autograd node (see _MlpFused)."""
    _check_cuda(x, w1, b1, w2, b2, resid)
    p = float(p) if training else 0.0
    seed1, seed2 = (_draw_seed(), _draw_seed()) if p > 0 else (0, 0)
    with torch.autocast("cuda", enabled=False):
        return _MlpFused.apply(x.to(torch.bfloat16).contiguous(), w1, b1, w2, b2,
                               None if resid is None else resid.contiguous(), p, seed1, seed2)


def linear_dropout_add(x, weight, bias, resid, p: float, training: bool):
    """``resid + dropout(linear(x, weight, bias), p)`` as one autograd node (``resid`` may be None)."""
    _check_cuda(x, weight, bias, resid)
    p = float(p) if training else 0.0
    dt = _autocast_dtype(x)
    if resid is not None:
        if resid.dtype not in (torch.float32, torch.bfloat16):
            resid = resid.to(torch.bfloat16)
        if resid.dtype == torch.bfloat16 and dt == torch.float32:
            resid = resid.float()
    if weight.shape[0] % 8:
        return dropout_add(linear(x, weight, bias), resid, p, training)
    seed = _draw_seed() if p > 0 else 0
    with torch.autocast("cuda", enabled=False):
        return _LinearDropoutAdd.apply(x.to(dt), weight, bias, None if resid is None else resid.contiguous(), p, seed)


def linear_gelu_dropout(x, weight, bias, p: float, training: bool):
    """``dropout(gelu(linear(x, weight, bias)), p)`` as one autograd node."""
    _check_cuda(x, weight, bias)
    p = float(p) if training else 0.0
    dt = _autocast_dtype(x)
    if weight.shape[0] % 8:
        return gelu_dropout(linear(x, weight, bias), p, training)
    seed = _draw_seed() if p > 0 else 0
    with torch.autocast("cuda", enabled=False):
        return _LinearGeluDropout.apply(x.to(dt), weight, bias, p, seed)


# ------------------------------------------------------------------------------------------------
# f4: token prologue - PatchEmbed + CLS + pos_embed + pos_drop (vit.py:25-36, 207-212)
# ------------------------------------------------------------------------------------------------
class _PatchEmbedTokens(torch.autograd.Function):
    """tokens = dropout(cat([cls, conv(img).flatten(2).T + b]) + pos): gvit_patchify, ONE library GEMM over all
    B*(1+Np) rows (the CLS slot of the patch matrix is zero), gvit_embed_assemble.  The image gets no gradient."""

    @staticmethod
    def forward(ctx, img, conv_w, conv_b, cls, pos, p, seed, dt, out_dt):
        B, C, H, W = img.shape
        D, P = conv_w.shape[0], conv_w.shape[-1]
        N, K = (H // P) * (W // P) + 1, C * P * P
        st = _stream()
        patches = torch.empty((B, N, K), dtype=dt, device=img.device)
        _call("gvit_patchify", _ptr(img), B, C, H, W, P, _dtype_code(img), _dtype_code(patches), _ptr(patches), st)
        y = _mm_nt(patches.view(B * N, K), _shadow(conv_w, dt).view(D, K))
        prm = [cls, pos] + ([conv_b] if conv_b is not None else [])
        if all(t.dtype == torch.float32 for t in prm) or all(t.dtype == dt for t in prm):
            pd = prm[0].dtype
        else:
            pd = dt
        if out_dt != dt and pd != torch.float32:
            out_dt = dt                             # an fp32 stream over a bf16 projection needs the fp32 parameters
        b_, c_, p_ = _shadow(conv_b, pd), _shadow(cls, pd), _shadow(pos, pd)
        out = torch.empty((B, N, D), dtype=out_dt, device=img.device)
        mask = torch.empty(B * N * D // 8, dtype=torch.uint8, device=img.device) if p > 0 else None
        _call("gvit_embed_assemble", _ptr(y), _ptr(b_), _ptr(c_), _ptr(p_), B, N, D, float(p), int(seed), 0,
              _rng_offset_ptr(), _dtype_code(y), GVIT_F32 if pd == torch.float32 else GVIT_BF16, _dtype_code(out), _ptr(out), _ptr(mask), st)
        ctx.save_for_backward(patches, mask)
        ctx.dt = dt
        ctx.p = p
        ctx.meta = (conv_w.shape, conv_w.dtype, None if conv_b is None else conv_b.dtype, cls.dtype, pos.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        patches, mask = ctx.saved_tensors
        w_shape, w_dt, b_dt, cls_dt, pos_dt = ctx.meta
        B, N, K = patches.shape
        D = dout.shape[-1]
        dout = dout.contiguous()
        if ctx.p > 0 or dout.dtype != ctx.dt:
            # keep mask (and, for an fp32 stream gradient, the cast to the GEMM dtype) in one pass
            d = torch.empty(dout.shape, dtype=ctx.dt, device=dout.device)
            _call("gvit_dropout_bwd", _ptr(dout), _ptr(mask), dout.numel(), float(ctx.p), _dtype_code(dout),
                  _dtype_code(d), _ptr(d), 0, 0, None, None, _stream())
        else:
            d = dout
        dsum = colsum(d.view(B, N * D)).view(N, D)              # fp32 sum over the batch: d pos_embed
        dpos = dsum.view(1, N, D).to(pos_dt) if ctx.needs_input_grad[4] else None
        dcls = dsum[0].view(1, 1, D).to(cls_dt) if ctx.needs_input_grad[3] else None
        dbias = dsum[1:].sum(0).to(b_dt) if (b_dt is not None and ctx.needs_input_grad[2]) else None
        dw = _wgrad(d.view(B * N, D), patches.view(B * N, K), w_dt).view(w_shape) if ctx.needs_input_grad[1] else None
        return None, dw, dbias, dcls, dpos, None, None, None, None


def patch_embed_supported(img: torch.Tensor, conv_w: torch.Tensor) -> bool:
    P = conv_w.shape[-1]
    return (img.dim() == 4 and conv_w.dim() == 4 and conv_w.shape[-2] == P and P % 8 == 0 and img.shape[1] == conv_w.shape[1]
            and img.shape[-1] % P == 0 and img.shape[-2] % P == 0 and conv_w.shape[0] % 8 == 0)


def patch_embed_tokens(img, conv_w, conv_b, cls_token, pos_embed, p: float, training: bool, fp32_stream: bool = False):
    """The (B, 1+Np, D) token tensor of vit.py:203-212 from the image: patch projection (Conv2d with kernel == stride,
    vit.py:22,34), CLS token, position embedding and pos_drop.  Runs in bf16 under autocast, else in fp32.
    ``fp32_stream``: under autocast emit the tokens in fp32 (bf16 projection + fp32 cls / pos: what torch.autocast's
    promotion rules give the reference at vit.py:207-211)."""
    _check_cuda(img, conv_w, conv_b, cls_token, pos_embed)
    if not patch_embed_supported(img, conv_w):
        raise ValueError(f"patch_embed_tokens needs a square patch size that is a multiple of 8 dividing the image; got image "
                         f"{tuple(img.shape)} and weight {tuple(conv_w.shape)}")
    p = float(p) if training else 0.0
    dt = torch.bfloat16 if torch.is_autocast_enabled("cuda") else _autocast_dtype(conv_w)
    if img.dtype not in (torch.float32, torch.bfloat16) or (img.dtype == torch.bfloat16 and dt == torch.float32):
        img = img.float()
    seed = _draw_seed() if p > 0 else 0
    with torch.autocast("cuda", enabled=False):
        return _PatchEmbedTokens.apply(img.contiguous(), conv_w, conv_b, cls_token, pos_embed, p, seed, dt,
                                       torch.float32 if fp32_stream else dt)


# ------------------------------------------------------------------------------------------------
# a7: graph construction
# ------------------------------------------------------------------------------------------------
def _token_view(h: torch.Tensor):
    """(pointer offset in bytes, batch stride, row stride, B, Np, D) of the patch tokens h[:, 1:, :]."""
    B, N, D = h.shape
    return D * h.element_size(), N * D, D, B, N - 1, D


@torch.no_grad()
def _knn(h: torch.Tensor, k: int):
    """G1-G3 dispatch on a contiguous (B, 1+Np, D) tensor.  bf16 with 256 < Np <= 1024 (the 576 tokens of a 384x384 image):
    Gram matrix on tcgen05 (gvit_bgemm, fp32) + row norms + per-row selection (gvit_knn_select); everything else is ONE
    gvit_knn_fwd launch (tcgen05 fused GEMM + top-k up to 256 tokens, the exact-fp32-FMA kernel for fp32 / odd shapes)."""
    off, bs, rs, B, Np, D = _token_view(h)
    st = _stream()
    idx = torch.empty((B, Np, k), dtype=torch.int32, device=h.device)
    vals = torch.empty((B, Np, k), dtype=torch.float32, device=h.device)
    rnorm = torch.empty((B, Np), dtype=torch.float32, device=h.device)
    if h.dtype == torch.bfloat16 and _lib.describe_path("knn", GVIT_BF16, Np, D).startswith("knn:tcgen05 gram"):
        tok = _ptr(h, off)
        ld = (Np + 3) // 4 * 4
        _call("gvit_dense_rownorm", tok, bs, rs, B, Np, D, _ptr(rnorm), st)
        G = torch.empty((B, Np, ld), dtype=torch.float32, device=h.device)
        _bgemm(B, Np, Np, [(tok, rs, bs, 0, tok, rs, bs, 0, D)], G, ld, Np * ld)
        _call("gvit_knn_select", _ptr(G), ld, _ptr(rnorm), B, Np, int(k), _ptr(idx), _ptr(vals), st)
    else:
        _call("gvit_knn_fwd", _ptr(h, off), bs, rs, B, Np, D, int(k), _dtype_code(h), _ptr(idx), _ptr(vals), _ptr(rnorm), st)
    return idx, vals, rnorm


@torch.no_grad()
def knn_graph(h: torch.Tensor, k: int):
    """G1-G3 over the patch tokens of h (B, 1+Np, D): returns (idx int32 (B,Np,k), vals fp32, rnorm fp32 (B,Np)).

    Neighbour order: descending cosine similarity, ties to the lowest index.
    """
    _check_cuda(h)
    return _knn(h.contiguous(), int(k))


@torch.no_grad()
def graph_reverse(idx: torch.Tensor):
    """Reverse adjacency (CSR) of idx (B,Np,k): (rev_ptr (B,Np+1), rev_src (B,Np*k)), ascending edge ids."""
    _check_cuda(idx)
    B, Np, k = idx.shape
    rev_ptr = torch.empty((B, Np + 1), dtype=torch.int32, device=idx.device)
    rev_src = torch.empty((B, Np * k), dtype=torch.int32, device=idx.device)
    _call("gvit_graph_reverse", _ptr(idx), B, Np, k, _ptr(rev_ptr), _ptr(rev_src), _stream())
    return rev_ptr, rev_src


@torch.no_grad()
def agg_gather(h: torch.Tensor, idx: torch.Tensor, vals: torch.Tensor):
    """G4+G5: w = softmax_k(vals), z_i = sum_j w_ij p[idx_ij]; returns (w fp32 (B,Np,k), z (B,Np,D))."""
    _check_cuda(h, idx, vals)
    h = h.contiguous()
    off, bs, rs, B, Np, D = _token_view(h)
    k = idx.shape[-1]
    w = torch.empty((B, Np, k), dtype=torch.float32, device=h.device)
    z = torch.empty((B, Np, D), dtype=h.dtype, device=h.device)
    _call("gvit_agg_gather_fwd", _ptr(h, off), bs, rs, B, Np, D, k, _dtype_code(h), _ptr(idx), _ptr(vals), _ptr(w),
          _ptr(z), _stream())
    return w, z


_FC1_ENABLED = {"on": True}
_FUSED_RESID_MAX_K = int(os.environ.get("GVIT_FUSED_RESID_MAX_K", "8192"))   # Linear + dropout + residual goes through the fused GEMM up to
# this reduction length (1024 while its main loop was L2-bound; the 2-SM main loop runs K = 3072 / 4096 at GEMM speed)


def fused_fc1_available(N: int, K: int) -> bool:
    return _FC1_ENABLED["on"] and _lib.describe_path("fc1", GVIT_BF16, N, K).startswith("fc1:tcgen05")


def fused_agg_available(dtype: torch.dtype, Np: int, D: int, k: int) -> bool:
    if dtype != torch.bfloat16 or k > 16:
        return False
    return _lib.describe_path("agg", GVIT_BF16, Np, D).startswith("agg:tcgen05")


def fused_agg_res32_available(Np: int, D: int, k: int) -> bool:
    """the fused aggregation kernel can add onto (and emit) an fp32 residual stream: the D <= 768 kernel only"""
    return k <= 16 and _lib.describe_path("agg_res32", GVIT_BF16, Np, D).startswith("agg:tcgen05")


def fused_graph_bwd_available(dtype: torch.dtype, Np: int, D: int, k: int) -> bool:
    if dtype != torch.bfloat16 or k > 16:
        return False
    return _lib.describe_path("graph_bwd", GVIT_BF16, Np, D).startswith("graph_bwd:tcgen05")


# ------------------------------------------------------------------------------------------------
# a7 + a8: the whole graph sub-layer as one differentiable operator
# ------------------------------------------------------------------------------------------------
class _PatchGraph(torch.autograd.Function):
    """y = cat([0, (softmax_k(S_knn) gathered p) Wg^T + b]) [+ resid], SURVEY.md section 9 G0-G6."""

    @staticmethod
    def forward(ctx, h, weight, bias, resid, k):
        ctx.w_dtype = weight.dtype
        ctx.b_dtype = bias.dtype if bias is not None else None
        weight, bias = _shadow(weight, h.dtype), _shadow(bias, h.dtype)
        off, bs, rs, B, Np, D = _token_view(h)
        dt = _dtype_code(h)
        st = _stream()
        idx, vals, rnorm = _knn(h, k)
        w = torch.empty((B, Np, k), dtype=torch.float32, device=h.device)
        if fused_agg_available(h.dtype, Np, D, k):
            out = torch.empty_like(h if resid is None else resid)        # an fp32 residual stream gets an fp32 result
            # z is laid out like h (zero CLS row) so that the weight / input gradients are plain GEMMs over all
            # B*(1+Np) rows of dout - slicing the CLS row off would cost a copy of dout per layer
            zf = torch.empty_like(h)
            zf[:, 0].zero_()
            _call("gvit_agg_fwd", _ptr(h), B, Np, D, k, dt, _ptr(idx), _ptr(vals), _ptr(weight), _ptr(bias),
                  _ptr(resid), _dtype_code(out), _ptr(out), _ptr(w), _ptr(zf, off), bs, st)
            z = zf
        else:
            z = torch.empty((B, Np, D), dtype=h.dtype, device=h.device)
            # fp32 parity path: fused gather + softmax kernel, then the projection as a library GEMM
            _call("gvit_agg_gather_fwd", _ptr(h, off), bs, rs, B, Np, D, k, dt, _ptr(idx), _ptr(vals), _ptr(w),
                  _ptr(z), st)
            out = torch.zeros_like(h) if resid is None else resid.clone()
            out[:, 1:] += F.linear(z, weight, bias)
        ctx.save_for_backward(h, weight, idx, rnorm, w, z, vals)
        ctx.k = k
        ctx.has = (bias is not None, resid is not None)
        ctx.mark_non_differentiable(idx, vals)
        return out, idx, vals

    @staticmethod
    def backward(ctx, dout, _didx, _dvals):
        h, weight, idx, rnorm, w, z, vals = ctx.saved_tensors
        k = ctx.k
        off, bs, rs, B, Np, D = _token_view(h)
        dt = _dtype_code(h)
        st = _stream()
        dout = dout.contiguous()
        dresid = dout if ctx.has[1] else None
        if z.shape[1] == Np + 1 and fused_graph_bwd_available(h.dtype, Np, D, k):
            # bf16: z / dz laid out like h; GEMMs over all rows (the CLS row of z is zero, the CLS row of dz is unused),
            # then both sparse stages as two per-image tensor-core GEMMs (no reverse adjacency)
            want_db = ctx.has[0] and ctx.needs_input_grad[2]
            if dout.dtype != h.dtype:
                # fp32 stream gradient: ONE pass casts it to the GEMM dtype and accumulates the column sums (bias gradient)
                d2 = torch.empty((B * (Np + 1), D), dtype=h.dtype, device=h.device)
                cs = torch.empty(D, dtype=torch.float32, device=h.device) if want_db else None
                ws = _colsum_ws(B * (Np + 1), D, h.device) if want_db else None
                _call("gvit_dropout_bwd", _ptr(dout), None, dout.numel(), 0.0, _dtype_code(dout), dt, _ptr(d2), D, Np + 1, _ptr(cs), _ptr(ws), st)
            else:
                d2 = dout.view(B * (Np + 1), D)
                cs = colsum(d2, skip_period=Np + 1) if want_db else None
            # the column sums leave the CLS rows out (skip_period): the projection only ever saw patch rows (G0), so its bias
            # gradient is exactly zero when they carry no gradient (the last block: the head reads the CLS token only)
            dweight = _wgrad(d2, z.view(B * (Np + 1), D), ctx.w_dtype) if ctx.needs_input_grad[1] else None
            dbias = cs.to(ctx.b_dtype) if want_db else None
            dh = None
            if ctx.needs_input_grad[0]:
                dz = _mm_nn(d2, weight)                            # (B*(1+Np), D)
                dvals = torch.empty((B, Np, k), dtype=torch.float32, device=h.device)
                dh = torch.empty_like(h)
                dh[:, 0].zero_()
                _call("gvit_graph_bwd", _ptr(h, off), bs, rs, B, Np, D, k, dt, _ptr(idx), _ptr(vals), _ptr(w),
                      _ptr(rnorm), _ptr(dz, off), bs, _ptr(dvals), _ptr(dh, off), st)
            return dh, dweight, dbias, dresid, None
        if z.shape[1] == Np + 1:
            z = z[:, 1:].contiguous()
        dy = dout[:, 1:, :]
        dy2 = dy.reshape(B * Np, D)
        dweight = _wgrad(dy2, z.view(B * Np, D), ctx.w_dtype) if ctx.needs_input_grad[1] else None
        dbias = dy2.float().sum(0).to(ctx.b_dtype) if (ctx.has[0] and ctx.needs_input_grad[2]) else None
        dh = None
        if ctx.needs_input_grad[0]:
            dz = (dy2 @ weight).view(B, Np, D)
            dvals = torch.empty((B, Np, k), dtype=torch.float32, device=h.device)
            dh = torch.empty_like(h)
            dh[:, 0].zero_()
            rev_ptr = torch.empty((B, Np + 1), dtype=torch.int32, device=h.device)
            rev_src = torch.empty((B, Np * k), dtype=torch.int32, device=h.device)
            _call("gvit_graph_reverse", _ptr(idx), B, Np, k, _ptr(rev_ptr), _ptr(rev_src), st)
            _call("gvit_agg_bwd", _ptr(h, off), bs, rs, B, Np, D, k, dt, _ptr(idx), _ptr(w), _ptr(dz), _ptr(rev_ptr),
                  _ptr(rev_src), _ptr(dvals), _ptr(dh, off), st)
            _call("gvit_knn_bwd", _ptr(h, off), bs, rs, B, Np, D, k, dt, _ptr(idx), _ptr(rnorm), _ptr(dvals),
                  _ptr(rev_ptr), _ptr(rev_src), _ptr(dh, off), st)
        return dh, dweight, dbias, dresid, None


# ------------------------------------------------------------------------------------------------
# a8, dense adjacency (graph_mode='dense', BASELINE configs[3]): batched tcgen05 GEMMs + row-wise libgvit kernels
# ------------------------------------------------------------------------------------------------
def _bgemm(batch, M, N, prods, out, out_rs, out_bs, row_scale=None):
    """out[b] = diag(row_scale[b]) sum_p op(A_p[b]) op(B_p[b]) through gvit_bgemm; prods = [(a_ptr, a_rs, a_bs, a_t, b_ptr, b_rs,
    b_bs, b_t, K), ...] (one or two products)."""
    p0 = prods[0]
    p1 = prods[1] if len(prods) > 1 else prods[0]
    _call("gvit_bgemm", batch, M, N, len(prods), *p0, *p1, _ptr(row_scale), _dtype_code(out), _ptr(out), out_rs, out_bs, _stream())


def dense_graph_available(dtype: torch.dtype, Np: int, D: int) -> bool:
    return dtype == torch.bfloat16 and _lib.describe_path("agg_dense", GVIT_BF16, Np, D).startswith("agg_dense:tcgen05")


class _DenseGraph(torch.autograd.Function):
    """y = cat([0, softmax(p^ p^^T) p Wg^T + b]) [+ resid] - SURVEY.md section 9 with G3 skipped (every patch token attends to
    every patch token of its image), bf16 storage / fp32 arithmetic, every matrix product a libgvit tcgen05 GEMM:

        forward   rn (G1) -> G = P P^T (fp32) -> A~ = softmax_j(G rn_i rn_j) (bf16) -> Z = A~ P -> out = resid + Z Wg^T + b
        backward  dZ = dY Wg -> dA~ = dZ P^T -> dG = A~ (dA~ - delta) rn rn^T -> T = A~^T dZ, V = (dG + dG^T) P
                  -> dp = T + V - rn^2 (p . V) p
    """

    @staticmethod
    def forward(ctx, h, weight, bias, resid):
        ctx.w_dtype = weight.dtype
        ctx.b_dtype = bias.dtype if bias is not None else None
        weight, bias = _shadow(weight, h.dtype), _shadow(bias, h.dtype)
        off, bs, rs, B, Np, D = _token_view(h)
        st = _stream()
        ld = (Np + 63) // 64 * 64
        tok = _ptr(h, off)
        rn = torch.empty((B, Np), dtype=torch.float32, device=h.device)
        _call("gvit_dense_rownorm", tok, bs, rs, B, Np, D, _ptr(rn), st)
        G = torch.empty((B, Np, ld), dtype=torch.float32, device=h.device)
        _bgemm(B, Np, Np, [(tok, rs, bs, 0, tok, rs, bs, 0, D)], G, ld, Np * ld)                    # G = P P^T
        A = torch.empty((B, Np, ld), dtype=h.dtype, device=h.device)
        _call("gvit_dense_softmax_fwd", _ptr(G), ld, _ptr(rn), B, Np, ld, _ptr(A), st)
        del G
        zf = torch.empty_like(h)                                      # laid out like h: weight / input gradients run over all rows
        zf[:, 0].zero_()
        _bgemm(B, Np, D, [(_ptr(A), ld, Np * ld, 0, tok, rs, bs, 1, Np)], zf[:, 1:], rs, bs)        # Z = A~ P
        M = B * (Np + 1)
        if resid is not None and D <= _FUSED_RESID_MAX_K and fused_fc1_available(D, D):
            out = torch.empty_like(resid)
            _call("gvit_linear_dropout_residual_fwd", _ptr(zf), _ptr(weight), _ptr(bias), _ptr(resid), M, D, D, 0.0, 0, 0, None,
                  GVIT_BF16, _dtype_code(resid), _ptr(out), None, st)
            out[:, 0] = resid[:, 0]                                   # the CLS row passes through untouched (G0): no bias either
        else:
            out = _mm_nt(zf, weight, bias)
            out[:, 0].zero_()
            if resid is not None:
                out = resid + out
        ctx.save_for_backward(h, weight, rn, A, zf)
        ctx.has = (bias is not None, resid is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        h, weight, rn, A, zf = ctx.saved_tensors
        off, bs, rs, B, Np, D = _token_view(h)
        dt, st = _dtype_code(h), _stream()
        ld = A.shape[-1]
        dout = dout.contiguous()
        dresid = dout if ctx.has[1] else None
        want_db = ctx.has[0] and ctx.needs_input_grad[2]
        rows = B * (Np + 1)
        if dout.dtype != h.dtype:                                     # fp32 stream gradient: one pass casts it and sums the columns
            d2 = torch.empty((rows, D), dtype=h.dtype, device=h.device)
            cs = torch.empty(D, dtype=torch.float32, device=h.device) if want_db else None
            ws = _colsum_ws(rows, D, h.device) if want_db else None
            _call("gvit_dropout_bwd", _ptr(dout), None, dout.numel(), 0.0, _dtype_code(dout), dt, _ptr(d2), D, Np + 1, _ptr(cs), _ptr(ws), st)
        else:
            d2 = dout.view(rows, D)
            cs = colsum(d2, skip_period=Np + 1) if want_db else None
        dweight = _wgrad(d2, zf.view(rows, D), ctx.w_dtype) if ctx.needs_input_grad[1] else None
        dbias = cs.to(ctx.b_dtype) if want_db else None
        dh = None
        if ctx.needs_input_grad[0]:
            tok = _ptr(h, off)
            dz = _mm_nn(d2, weight)                                       # (B*(1+Np), D); its CLS rows are never read
            dzt = _ptr(dz, off)
            dA = torch.empty((B, Np, ld), dtype=torch.float32, device=h.device)
            _bgemm(B, Np, Np, [(dzt, rs, bs, 0, tok, rs, bs, 0, D)], dA, ld, Np * ld)               # dA~ = dZ P^T
            dG = torch.empty((B, Np, ld), dtype=h.dtype, device=h.device)
            _call("gvit_dense_softmax_bwd", _ptr(dA), ld, _ptr(A), ld, _ptr(rn), B, Np, _ptr(dG), st)
            del dA
            V = torch.empty((B, Np, D), dtype=h.dtype, device=h.device)
            _bgemm(B, Np, D, [(_ptr(dG), ld, Np * ld, 0, tok, rs, bs, 1, Np),                       # V = dG P + dG^T P
                              (_ptr(dG), ld, Np * ld, 1, tok, rs, bs, 1, Np)], V, D, Np * D)
            T = torch.empty((B, Np, D), dtype=h.dtype, device=h.device)
            _bgemm(B, Np, D, [(_ptr(A), ld, Np * ld, 1, dzt, rs, bs, 1, Np)], T, D, Np * D)         # T = A~^T dZ
            dh = torch.empty_like(h)
            dh[:, 0].zero_()
            _call("gvit_dense_combine_bwd", _ptr(T), _ptr(V), tok, bs, rs, _ptr(rn), B, Np, D, _ptr(dh, off), st)
        return dh, dweight, dbias, dresid


def dense_graph(h: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, resid: torch.Tensor | None = None):
    """The dense-adjacency graph sub-layer on layer-normed tokens h (B, 1+Np, D) under bf16 autocast; returns (B, 1+Np, D) in
    the dtype of ``resid`` (or of the compute dtype).  The CLS row is zero / ``resid``'s CLS row."""
    _check_cuda(h, weight, bias, resid)
    if h.dim() != 3 or h.shape[1] < 2:
        raise ValueError(f"h must be (B, 1+Np, D); got {tuple(h.shape)}")
    dt = _autocast_dtype(h)
    if not dense_graph_available(dt, h.shape[1] - 1, h.shape[2]):
        raise _lib.GvitError("dense_graph", 5, "the native dense graph layer needs bf16 compute, Np <= 1024, D % 64 == 0, D <= 1024")
    with torch.autocast("cuda", enabled=False):
        return _DenseGraph.apply(h.to(dt).contiguous(), weight, bias, None if resid is None else resid.contiguous())


def patch_graph(h: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, k: int,
                resid: torch.Tensor | None = None, return_graph: bool = False):
    """The kNN graph sub-layer on layer-normed tokens h (B, 1+Np, D): returns (B, 1+Np, D).

    The CLS row of the result is zero (or ``resid``'s CLS row when a residual is fused in).
    ``return_graph=True`` also returns the adjacency (idx int32, vals fp32) that was built.
    """
    _check_cuda(h, weight, bias, resid)
    if h.dim() != 3 or h.shape[1] < 2:
        raise ValueError(f"h must be (B, 1+Np, D); got {tuple(h.shape)}")
    dt = _autocast_dtype(h)
    with torch.autocast("cuda", enabled=False):
        B_, N_, D_ = h.shape
        fuse_resid = resid is not None and (resid.dtype == dt or (
            resid.dtype == torch.float32 and dt == torch.bfloat16 and fused_agg_res32_available(N_ - 1, D_, int(k))))
        out, idx, vals = _PatchGraph.apply(h.to(dt).contiguous(), weight, bias,
                                           resid.contiguous() if fuse_resid else None, int(k))
        if resid is not None and not fuse_resid:
            out = resid + out
    return (out, idx, vals) if return_graph else out
