"""TEST INFRASTRUCTURE ONLY - CPU oracle for the patch-token graph sub-layer.

**Parity unpinned**: the reference repository has no graph code (SURVEY.md
section 0), so this file *is* the specification (SURVEY.md section 9, steps
G0-G6, frozen as ``oracle.GRAPH_SPEC_VERSION``).  It is written in the idiom of
the reference's layers so it reads as "the reference's graph layer":

* pre-norm residual sub-layer wrapped by ``drop_path`` exactly like
  ``Block.forward`` (/root/reference/src/models/vit.py:116-119);
* ``nn.Linear`` projection initialised by the container's ``_init_weights``
  (/root/reference/src/models/vit.py:173-180);
* CLS token at index 0 (/root/reference/src/models/vit.py:207-208,219) - the
  graph is built over the Np patch tokens only and leaves CLS untouched.

Plain PyTorch, fp32 on CPU, autograd supplies the backward.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

NORM_EPS = 1e-12  # F.normalize default; G1


def l2_normalize(p: torch.Tensor) -> torch.Tensor:
    """G1: p_hat = p / max(||p||_2, 1e-12), always in fp32."""
    p = p.float()
    return p / p.norm(dim=-1, keepdim=True).clamp_min(NORM_EPS)


def similarity(p: torch.Tensor) -> torch.Tensor:
    """G2: S = p_hat p_hat^T, (B,Np,Np), fp32 accumulate; self-similarity kept."""
    ph = l2_normalize(p)
    return ph @ ph.transpose(-1, -2)


def knn_select(S: torch.Tensor, k: int):
    """G3: per-row k largest, descending, ties -> lowest index (stable sort).

    ``torch.topk`` is *not* used: it does not break ties by lowest index
    ("Hard parts").
    """
    order = torch.sort(S, dim=-1, descending=True, stable=True).indices
    idx = order[..., :k]
    return idx, S.gather(-1, idx)


def knn_select_strict(p: torch.Tensor, k: int):
    """G1-G3 in the STRICT fp32 order of GRAPH_SPEC_VERSION 2 (``oracle/knn_strict.c``: sequential fp32 FMA chains,
    IEEE division / sqrt, ties -> lowest index).  This is the statement "kNN neighbour indices bit-exact in fp32" is
    checked against: returns (idx int64 (B,Np,k), vals fp32 (B,Np,k)) as torch tensors.  ``similarity`` + ``knn_select``
    above are the same mathematics with the GEMM's (MKL's) unspecified accumulation order; the two agree on every row
    whose decision margin exceeds fp32 round-off (tests/test_oracle_golden.py)."""
    from . import knn_strict
    idx, vals, _ = knn_strict.knn_strict(p.detach().float().cpu().numpy(), k)
    return torch.from_numpy(idx.astype(np.int64)), torch.from_numpy(vals)


def knn_f64(p: np.ndarray, k: int):
    """Float64 numpy restatement of G1-G3 for index parity at scale.

    Returns (idx int32 (B,Np,k), vals float64, margin float64 (B,Np)) where
    ``margin`` is the smallest gap between consecutive kept similarities and
    between the k-th kept and the best rejected one: rows whose margin exceeds
    the fp32 accumulation noise (~1e-5) have an unambiguous answer, so a kernel
    must reproduce their indices bit-exactly; exact ties (margin 0 from
    duplicated rows) must resolve to the lowest index.
    """
    p = np.asarray(p, dtype=np.float64)
    n = np.maximum(np.linalg.norm(p, axis=-1, keepdims=True), NORM_EPS)
    ph = p / n
    S = ph @ np.swapaxes(ph, -1, -2)
    order = np.argsort(-S, axis=-1, kind="stable")
    idx = order[..., :k]
    srt = np.take_along_axis(S, order[..., : k + 1], axis=-1)
    gaps = srt[..., :-1] - srt[..., 1:] if srt.shape[-1] > 1 else np.full(S.shape[:-1] + (1,), np.inf)
    return idx.astype(np.int32), np.take_along_axis(S, idx, axis=-1), gaps.min(axis=-1)


def graph_layer_forward(h, weight, bias, k=8, mode="knn", compute_dtype=None, return_aux=False, idx_override=None,
                        strict=True):
    """G0-G6 on an already layer-normed token tensor h (B, 1+Np, D).

    compute_dtype: None -> everything fp32.  torch.bfloat16 emulates autocast:
    G1-G4 stay fp32, G5-G6 run on operands rounded to compute_dtype (matmuls
    under autocast cast both operands; accumulation is fp32).
    strict (knn mode, no idx_override): the neighbour INDICES come from the strict-order fp32 evaluation
    (``knn_select_strict``, spec version 2); the similarity VALUES that carry the gradient stay torch's (autograd).
    """
    B, N, D = h.shape
    p = h[:, 1:, :]                                   # G0
    S = similarity(p)                                 # G1, G2
    cd = compute_dtype or torch.float32
    if mode == "knn":
        if idx_override is not None:
            idx = idx_override.long()
        elif strict:
            idx, _ = knn_select_strict(p, k)          # G3, strict fp32 order
            idx = idx.to(h.device)
        else:
            idx, _ = knn_select(S, k)                 # G3, accumulation order of the library GEMM
        vals = S.gather(-1, idx)
        w = torch.softmax(vals, dim=-1)               # G4
        pg = p.to(cd)
        bi = torch.arange(B, device=h.device)[:, None, None]
        nb = pg[bi, idx]                              # (B,Np,k,D) gather of un-normalised tokens
        z = (w.to(cd).unsqueeze(-1) * nb).float().sum(dim=2).to(cd)   # G5
    elif mode == "dense":
        idx, vals = None, S
        w = torch.softmax(S, dim=-1)                  # G4 over all Np
        z = (w.to(cd) @ p.to(cd))                     # G5
    else:
        raise ValueError(f"unknown graph mode {mode!r}")
    y = F.linear(z, weight.to(cd), None if bias is None else bias.to(cd))   # G6
    out = torch.cat([torch.zeros(B, 1, D, dtype=y.dtype, device=y.device), y], dim=1)
    if return_aux:
        return out, {"idx": idx, "vals": vals, "w": w, "z": z, "S": S}
    return out


class PatchGraphLayer(nn.Module):
    """``PatchGraphLayer(dim, k=8, mode='knn')``; forward(h:(B,1+Np,D)) -> (B,1+Np,D)."""

    def __init__(self, dim, k=8, mode="knn"):
        super().__init__()
        self.k, self.mode = k, mode
        self.proj = nn.Linear(dim, dim)
        # tests may pin the adjacency (e.g. to the one a device run built from ITS tokens of this layer): deep in a network
        # a near-tie can flip on 1e-7 of input noise, and a flipped neighbour is not a small perturbation of the output
        self.idx_override = None

    def forward(self, h):
        return graph_layer_forward(h, self.proj.weight, self.proj.bias, self.k, self.mode, idx_override=self.idx_override)
