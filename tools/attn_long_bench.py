"""Time gvit_attn_fwd / gvit_attn_bwd at a long sequence (default: ViT-L/16 at 384x384, B = 32, N = 577, H = 16) and check the
forward against an fp32 softmax(QK^T)V of the same bf16 inputs.  python tools/attn_long_bench.py [--batch 32] [--tokens 577] [--heads 16]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_augmented_vision_transformers_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--tokens", type=int, default=577)
ap.add_argument("--heads", type=int, default=16)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, N, H, dh = a.batch, a.tokens, a.heads, 64
g = torch.Generator(device=dev).manual_seed(0)
qkvs = [torch.randn(B, N, 3 * H * dh, device=dev, dtype=torch.bfloat16, generator=g).requires_grad_(True) for _ in range(3)]
scale = dh ** -0.5
out = ops.attention_core(qkvs[0], H, scale)
q, k, v = qkvs[0].detach()[:2].float().view(2, N, 3, H, dh).permute(2, 0, 3, 1, 4)
want = (torch.softmax(q @ k.transpose(-1, -2) * scale, -1) @ v).transpose(1, 2).reshape(2, N, H * dh)
err = float((out.detach()[:2].float() - want).abs().max() / want.abs().max())
cot = torch.randn_like(out)


def timeit(fn):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters


with torch.no_grad():
    t_fwd = timeit(lambda i: ops.attention_core(qkvs[i % 3], H, scale))


def fb(i):
    x = qkvs[i % 3]
    x.grad = None
    ops.attention_core(x, H, scale).backward(cot)


t_fb = timeit(fb)
f_fwd = 4.0 * B * N * N * H * dh
print(json.dumps({"B": B, "N": N, "H": H, "fwd_rel_err": err, "fwd_ms": round(t_fwd, 4), "fwd_tflops": round(f_fwd / t_fwd / 1e9, 1),
                  "bwd_ms": round(t_fb - t_fwd, 4), "bwd_tflops": round(2.5 * f_fwd / (t_fb - t_fwd) / 1e9, 1)}))
