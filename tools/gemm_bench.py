"""gvit_linear_gemm (2-SM tcgen05 GEMM, csrc/gemm2_tc.cu) against torch.matmul (cuBLASLt) on the same box, same shapes:
correctness (max relative error against an fp32 product of the same bf16 operands) and time (CUDA events, inputs
rotated through > L2 worth of buffers).  python tools/gemm_bench.py [--batch 256] [--iters 20] [--quick]
Prints one JSON line per (layer, product)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_augmented_vision_transformers_b200 import _lib  # noqa: E402
from graph_augmented_vision_transformers_b200._lib import GVIT_BF16, GVIT_F32  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--tokens", type=int, default=197)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--quick", action="store_true", help="correctness only, small M")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
_lib.load()
st = torch.cuda.current_stream().cuda_stream
ws = torch.empty(64 << 20, dtype=torch.float32, device=dev)                 # split-K partial tiles (256 MB: any shape here)


def gemm(a, a_t, b, b_t, M, N, K, bias, out):
    _lib.call("gvit_linear_gemm", a.data_ptr(), a_t, a.stride(0), b.data_ptr(), b_t, b.stride(0), M, N, K,
              bias.data_ptr() if bias is not None else None, GVIT_F32 if out.dtype == torch.float32 else GVIT_BF16,
              out.data_ptr(), out.stride(0), ws.data_ptr(), ws.numel() * 4, st)


def timeit(fn, nsets, iters):
    for i in range(3):
        fn(i % nsets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nsets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def relerr(got, want):
    return float((got.float() - want).abs().max() / want.abs().max().clamp_min(1e-30))


M = args.batch * args.tokens if not args.quick else 1000
D = args.dim
layers = [("qkv", D, 3 * D), ("proj", D, D), ("fc1", D, 4 * D), ("fc2", 4 * D, D)]
g = torch.Generator(device=dev).manual_seed(0)
for name, K, N in layers:
    per_set = 2 * (M * K + M * N + N * K)
    nsets = max(2, min(8, int(300e6 // per_set) + 1))
    xs = [torch.randn(M, K, device=dev, generator=g).bfloat16() for _ in range(nsets)]
    dys = [torch.randn(M, N, device=dev, generator=g).bfloat16() for _ in range(nsets)]
    w = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev, generator=g).bfloat16()
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
    dw = torch.empty(N, K, device=dev, dtype=torch.float32)
    flops = 2.0 * M * N * K
    prods = {
        "fwd": (lambda i: gemm(xs[i], 0, w, 0, M, N, K, bias, y), lambda i: torch.nn.functional.linear(xs[i], w, bias),
                lambda: (y, xs[0].float() @ w.float().t() + bias.float())),
        "dgrad": (lambda i: gemm(dys[i], 0, w, 1, M, K, N, None, dx), lambda i: torch.mm(dys[i], w),
                  lambda: (dx, dys[0].float() @ w.float())),
        "wgrad": (lambda i: gemm(dys[i], 1, xs[i], 1, N, K, M, None, dw), lambda i: torch.mm(dys[i].t(), xs[i], out_dtype=torch.float32),
                  lambda: (dw, dys[0].float().t() @ xs[0].float())),
    }
    for pname, (ours, lib, check) in prods.items():
        if pname == "dgrad" and K % 256:
            continue
        if pname == "wgrad" and K % 256:
            continue
        ours(0)
        torch.cuda.synchronize()
        got, want = check()
        err = relerr(got, want)
        rec = {"layer": name, "product": pname, "M": M, "N": N, "K": K, "rel_err": err}
        if pname == "wgrad":
            ours(0)                                     # run-to-run determinism of the split-K order
            first = dw.clone()
            ours(0)
            torch.cuda.synchronize()
            rec["deterministic"] = bool(torch.equal(first, dw))
            rec["ws_bytes"] = _lib.load().gvit_linear_gemm_ws_bytes(N, K, M)
        if not args.quick:
            t_ours, t_lib = timeit(ours, nsets, args.iters), timeit(lib, nsets, args.iters)
            rec.update(ms=round(t_ours, 4), tflops=round(flops / t_ours / 1e9, 1), lib_ms=round(t_lib, 4),
                       lib_tflops=round(flops / t_lib / 1e9, 1))
        print(json.dumps(rec), flush=True)
