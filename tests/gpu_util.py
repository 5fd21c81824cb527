"""Shared helpers of the -m gpu parity tests (every comparison goes through the C ABI via the ops/modules)."""
import numpy as np
import torch

from oracle import graph_oracle

DEV = "cuda"


def tokens(B, Np, D, seed=0, dtype=torch.float32):
    """Layer-norm-like tokens (B, 1+Np, D), rounded to dtype; returned as fp32 CPU master + device copy."""
    g = torch.Generator().manual_seed(seed)
    h = torch.randn(B, Np + 1, D, generator=g)
    h = (h - h.mean(-1, keepdim=True)) / h.std(-1, keepdim=True)
    h = h.to(dtype)
    return h.float(), h.to(DEV)


def check_adjacency(h_cpu, idx_dev, vals_dev, k, noise, min_sure=None):
    """idx must equal the float64 oracle's on every row whose decision margin exceeds the arithmetic noise;
    on the remaining rows the selected similarities must still agree (two near-equal neighbours may swap)."""
    idx64, vals64, margin = graph_oracle.knn_f64(h_cpu[:, 1:].numpy(), k)
    idx = idx_dev.cpu().numpy()
    vals = vals_dev.cpu().numpy()
    sure = margin > noise
    if min_sure is None:                      # k+1 gaps per row can each fall inside the noise band
        min_sure = 0.9 if k <= 16 else 0.7
    assert sure.mean() >= min_sure, f"only {sure.mean():.3f} of the rows are decidable"
    bad = (idx[sure] != idx64[sure]).any(-1)
    assert not bad.any(), f"{bad.sum()} decidable rows differ, e.g. {idx[sure][bad][:2]} vs {idx64[sure][bad][:2]}"
    assert np.abs(vals - vals64).max() < max(noise, 2e-6), np.abs(vals - vals64).max()
    assert (np.diff(vals, axis=-1) <= 0).all(), "neighbours are not in descending-similarity order"
    return sure
