"""Kernel-time table of one bench training step (torch.profiler / CUPTI): which kernels the 74 ms go to.
Usage (GPU box): python tools/profile_step.py [--batch 256] > gpurun_out/step_profile.txt"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from graph_augmented_vision_transformers_b200 import modules  # noqa: E402
from graph_augmented_vision_transformers_b200.losses import DynamicWeightedLoss  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--config", choices=["vitb224", "vitl384"], default="vitb224")
ap.add_argument("--rows", type=int, default=60)
ap.add_argument("--ncu", action="store_true", help="no torch profiler: bracket one step with cudaProfilerStart/Stop for ncu --profile-from-start off")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(42)
cfg, dbatch, _, _ = bench.WORKLOADS[args.config]
args.batch = args.batch or dbatch
model = modules.VisionTransformer(**cfg).to(dev).train()
crit = DynamicWeightedLoss(14).to(dev)
from graph_augmented_vision_transformers_b200 import optim  # noqa: E402
opt = optim.FusedAdamW([{"params": model.parameters()}, {"params": crit.parameters(), "lr": 1e-5}], lr=1e-4, weight_decay=0.05,
                       max_norm=1.0, warmup_steps=100, total_steps=10000)
img = torch.randn(args.batch, 3, cfg["img_size"], cfg["img_size"], device=dev)
tgt = (torch.rand(args.batch, 14, device=dev) > 0.9).float()
params = list(model.parameters()) + list(crit.parameters())


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(img)
    loss, _ = crit(logits, tgt)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
if args.ncu:
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
evts = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
evts.sort(key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in evts)
print(f"total device kernel time {total / 1e3:.2f} ms over {sum(e.count for e in evts)} launches")
for e in evts[: args.rows]:
    print(f"{e.device_time_total / 1e3:9.3f} ms {100 * e.device_time_total / total:5.1f}% x{e.count:<5d} {e.key[:150]}")
