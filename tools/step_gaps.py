"""How much of a replayed (CUDA-graph) training step is NOT inside a kernel: CUPTI timeline of one replay -> span from the first
kernel's start to the last kernel's end, summed kernel time, and the distribution of the gaps between consecutive kernels.
python tools/step_gaps.py [--batch 256]"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from graph_augmented_vision_transformers_b200 import modules, optim, step  # noqa: E402
from graph_augmented_vision_transformers_b200.losses import DynamicWeightedLoss  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
cfg, _, _, _ = bench.WORKLOADS["vitb224"]
model = modules.VisionTransformer(**cfg).to(dev).train()
crit = DynamicWeightedLoss(14).to(dev)
opt = optim.FusedAdamW([{"params": model.parameters()}, {"params": crit.parameters(), "lr": 1e-5}], lr=1e-4, weight_decay=0.05,
                       max_norm=1.0, warmup_steps=100, total_steps=10000)
img = torch.randn(a.batch, 3, cfg["img_size"], cfg["img_size"], device=dev)
tgt = (torch.rand(a.batch, 14, device=dev) > 0.9).float()
cap = step.CapturedTrainStep(model, crit, opt, max_norm=None)
for _ in range(3):
    cap(img, tgt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    cap(img, tgt)
e1.record()
torch.cuda.synchronize()
print(f"replayed step: {e0.elapsed_time(e1) / 5:.3f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    cap(img, tgt)
    torch.cuda.synchronize()
ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
            key=lambda t: t[0])
span = (ev[-1][1] - ev[0][0]) / 1e3
busy = sum(e - s for s, e, _ in ev) / 1e3
gaps = [max(0.0, ev[i + 1][0] - max(x[1] for x in ev[max(0, i - 3): i + 1])) for i in range(len(ev) - 1)]
big = sorted(range(len(gaps)), key=lambda i: -gaps[i])[:6]
for i in big:
    print(f"  gap {gaps[i]:9.2f} us after [{ev[i][2][:70]}] ({ev[i][1] - ev[i][0]:.1f} us) before [{ev[i + 1][2][:70]}]")
gaps.sort()
n = len(gaps)
print(f"{len(ev)} kernels, span {span:.3f} ms, summed kernel time {busy:.3f} ms, idle {span - busy:.3f} ms ({100 * (span - busy) / span:.1f} %)")
print(f"gap between consecutive kernels (us): median {gaps[n // 2]:.2f}, p90 {gaps[int(n * 0.9)]:.2f}, max {gaps[-1]:.2f}, sum {sum(gaps) / 1e3:.3f} ms")

# ---- three replays back to back: what separates one replay's last kernel from the next replay's first (steady state) ----
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        cap(img, tgt)
    torch.cuda.synchronize()
ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
            key=lambda t: t[0])
gaps = sorted(((ev[i + 1][0] - ev[i][1], i) for i in range(len(ev) - 1)), reverse=True)[:4]
print(f"three replays back to back: {len(ev)} device events, total span {(ev[-1][1] - ev[0][0]) / 1e3:.3f} ms")
for g, i in gaps:
    print(f"  gap {g:9.2f} us after [{ev[i][2][:60]}] before [{ev[i + 1][2][:60]}] (event {i + 1} of {len(ev)})")
