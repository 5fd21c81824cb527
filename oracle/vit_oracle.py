"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference ViT hot path.

Every function cites the line of /root/reference/src/models/vit.py it follows.
State-dict keys are identical to the reference's, so weights move freely
between the reference, this oracle and the CUDA-backed modules.  Pinned by
``tests/test_oracle_vs_reference.py`` (live import of the reference when
/root/reference is present) and by the fixtures in ``tests/golden/`` that
``oracle/make_golden.py`` produced from the imported reference.

``graph_mode=None`` reproduces the reference model exactly; with
``graph_mode in {'knn','dense'}`` block i gains the SURVEY.md section 9
sub-layer when ``i % graph_every == 0``.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .graph_oracle import PatchGraphLayer


def attention_core(qkv: torch.Tensor, num_heads: int, scale: float) -> torch.Tensor:
    """vit.py:59-69 - qkv is the packed (B,N,3*C) projection output.

    reshape (B,N,3,H,dh) -> per-head q,k,v; softmax((q k^T) * scale) v; heads are
    concatenated back to (B,N,C) in head-major order.
    """
    B, N, C3 = qkv.shape
    dh = C3 // 3 // num_heads
    q, k, v = qkv.view(B, N, 3, num_heads, dh).permute(2, 0, 3, 1, 4)      # vit.py:59-61
    s = torch.matmul(q, k.transpose(-1, -2)) * scale                        # vit.py:64
    p = torch.softmax(s, dim=-1)                                            # vit.py:65
    return torch.matmul(p, v).transpose(1, 2).reshape(B, N, C3 // 3)        # vit.py:69


def attention_forward(x, w_qkv, b_qkv, w_proj, b_proj, num_heads, p_drop=0.0, training=False):
    """vit.py:55-72 (attn_drop is 0 in every shipped config and is omitted)."""
    C = x.shape[-1]
    scale = (C // num_heads) ** -0.5                                        # vit.py:46-47
    o = attention_core(F.linear(x, w_qkv, b_qkv), num_heads, scale)
    return F.dropout(F.linear(o, w_proj, b_proj), p_drop, training)        # vit.py:70-71


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        if dim % num_heads:
            raise AssertionError("dim should be divisible by num_heads")    # vit.py:44
        self.num_heads, self.scale = num_heads, (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, 3 * dim, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        o = attention_core(self.qkv(x), self.num_heads, self.scale)
        return self.proj_drop(self.proj(o))


class Mlp(nn.Module):
    """vit.py:75-94 - fc1, exact-erf GELU, dropout, fc2, dropout."""

    def __init__(self, in_features, hidden_features=None, out_features=None, drop=0.0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


def drop_path(x, p, training):
    """vit.py:234-242 - per-sample stochastic depth."""
    if p == 0.0 or not training:
        return x
    keep = 1.0 - p
    mask = torch.floor(keep + torch.rand((x.shape[0],) + (1,) * (x.ndim - 1), dtype=x.dtype, device=x.device))
    return x.div(keep) * mask


class Block(nn.Module):
    """vit.py:97-119 plus the optional section-9 graph sub-layer between the two lines."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, *, graph_mode=None, graph_k=8):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = Attention(dim, num_heads, qkv_bias, attn_drop, drop)
        if graph_mode is not None:
            self.norm_g = nn.LayerNorm(dim)
            self.graph = PatchGraphLayer(dim, graph_k, graph_mode)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), drop=drop)
        self.dp = float(drop_path)
        self.graph_mode = graph_mode

    def forward(self, x):
        x = x + drop_path(self.attn(self.norm1(x)), self.dp, self.training)       # vit.py:117
        if self.graph_mode is not None:
            x = x + drop_path(self.graph(self.norm_g(x)), self.dp, self.training)
        return x + drop_path(self.mlp(self.norm2(x)), self.dp, self.training)      # vit.py:118


class PatchEmbed(nn.Module):
    """vit.py:12-36."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size, self.num_patches = img_size, (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, patch_size, patch_size)

    def forward(self, x):
        if x.shape[-2] != self.img_size or x.shape[-1] != self.img_size:         # vit.py:27-28
            raise AssertionError(f"Input image size ({x.shape[-2]}*{x.shape[-1]}) doesn't match")
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    """vit.py:122-224 with keyword-only graph extensions."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=14, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.0, *, graph_mode=None, graph_k=8, graph_every=1):
        super().__init__()
        self.num_classes, self.num_features = num_classes, embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.pos_drop = nn.Dropout(drop_rate)
        rates = torch.linspace(0, drop_path_rate, depth).tolist()                # vit.py:144
        self.blocks = nn.ModuleList(
            Block(embed_dim, num_heads, mlp_ratio, qkv_bias, drop_rate, attn_drop_rate, rates[i],
                  graph_mode=graph_mode if (graph_mode is not None and i % graph_every == 0) else None,
                  graph_k=graph_k)
            for i in range(depth))
        self.norm = nn.LayerNorm(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        self._reset()

    def _reset(self):                                                            # vit.py:162-180
        w = self.patch_embed.proj.weight.data
        nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.zeros_(m.bias)
                nn.init.ones_(m.weight)

    def forward_features(self, x):                                               # vit.py:202-219
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1) + self.pos_embed
        x = self.pos_drop(x)
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x)[:, 0]

    def forward(self, x):                                                        # vit.py:221-224
        return self.head(self.forward_features(x))


def multilabel_loss(logits, targets, lambdas, pos_weight, gamma=2.0):
    """/root/reference/src/training/losses.py:26-68 - three-term learnable-weighted loss."""
    w = torch.softmax(lambdas, dim=0)                                            # losses.py:28-32
    wbce = F.binary_cross_entropy_with_logits(logits, targets, pos_weight=pos_weight)   # :35-37
    bce = F.binary_cross_entropy_with_logits(logits, targets, reduction="none")         # :40-42
    focal = ((1 - torch.exp(-bce)) ** gamma * bce).mean()                        # :43-44
    sp = torch.sigmoid(logits)                                                   # :47
    pos = targets * torch.log(sp.clamp(min=1e-8)) * (1 - sp)                     # :51, gamma_pos=1
    neg = (1 - targets) * torch.log((1 - sp).clamp(min=1e-8)) * sp.pow(4)        # :52, gamma_neg=4
    asl = -(pos + neg).mean()                                                    # :53
    return w[0] * wbce + w[1] * focal + w[2] * asl                               # :56-60
