"""Where the data-parallel step loses time against the single-GPU step: kernel timeline of ONE eager training step on
rank 0 (torch.profiler / CUPTI), split into compute kernels and NCCL kernels.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/dp_overlap.py > gpurun_out/dp_overlap_nN.txt

Reported: total NCCL kernel time, the part of it during which NO compute kernel was running on the device (exposed
communication), the compute kernels' busy time, and the same step without gradient sync for comparison."""
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from graph_augmented_vision_transformers_b200 import dp, modules  # noqa: E402
from graph_augmented_vision_transformers_b200.losses import DynamicWeightedLoss  # noqa: E402

rank, world, local = dp.init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
torch.manual_seed(42)
model = modules.VisionTransformer(**bench.MODEL_CFG).to(dev).train()
crit = DynamicWeightedLoss(14).to(dev)
dp.broadcast_parameters(model)
dp.broadcast_parameters(crit)
sync = dp.GradSync(model, bucket_mb=32.0, extra_params=list(crit.parameters()))
params = list(model.parameters()) + list(crit.parameters())
opt = torch.optim.AdamW([{"params": model.parameters()}, {"params": crit.parameters(), "lr": 1e-5}], lr=1e-4, weight_decay=0.05,
                        fused=True)
B = bench.PER_GPU_BATCH
img = torch.randn(B, 3, 224, 224, device=dev)
tgt = (torch.rand(B, 14, device=dev) > 0.9).float()


def step(with_sync=True):
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(img)
    loss, _ = crit(logits, tgt)
    loss.backward()
    if with_sync:
        sync.finish()
    torch.nn.utils.clip_grad_norm_(params, 1.0, foreach=True)
    opt.step()


def union(iv):
    iv = sorted(iv)
    out = []
    for a, b in iv:
        if out and a <= out[-1][1]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return out


def length(iv):
    return sum(b - a for a, b in iv)


def subtract(a_iv, b_iv):
    """total length of a_iv not covered by b_iv (both unions)"""
    tot, j = 0.0, 0
    for a, b in a_iv:
        cur = a
        while j < len(b_iv) and b_iv[j][1] <= cur:
            j += 1
        k = j
        while k < len(b_iv) and b_iv[k][0] < b:
            if b_iv[k][0] > cur:
                tot += b_iv[k][0] - cur
            cur = max(cur, b_iv[k][1])
            k += 1
        if cur < b:
            tot += b - cur
    return tot


for _ in range(4):
    step()
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0]
    nccl = [(e.time_range.start, e.time_range.end) for e in ev if "nccl" in e.name.lower()]
    comp = [(e.time_range.start, e.time_range.end) for e in ev if "nccl" not in e.name.lower()]
    un, uc = union(nccl), union(comp)
    t0, t1 = min(a for a, _ in nccl + comp), max(b for _, b in nccl + comp)
    print(f"# world {world}, rank 0, one eager training step of bench.py's workload (B = {B} per rank), times in ms")
    print(f"step span (first kernel start .. last kernel end)   {(t1 - t0) / 1e3:9.3f}")
    print(f"compute kernels busy (union)                        {length(uc) / 1e3:9.3f}")
    print(f"NCCL kernels: {len(nccl)} launches, busy (union)            {length(un) / 1e3:9.3f}")
    print(f"NCCL time with NO compute kernel running (exposed)  {subtract(un, uc) / 1e3:9.3f}")
    print(f"buckets {len(sync.buckets)}: sizes MB {[round(sum(p.numel() * 4 for p in b) / 2**20, 1) for b in sync.buckets]}")
    for a, b in nccl:
        print(f"  nccl kernel at {(a - t0) / 1e3:8.3f} .. {(b - t0) / 1e3:8.3f}  ({(b - a) / 1e3:6.3f} ms)")
# the same step with and without gradient sync (hooks removed: what a single GPU does), CUDA-event timed
for w in (True, False):
    if not w:
        sync.remove()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        step(with_sync=w)
    b.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"eager step, gradient sync {'ON ' if w else 'OFF'}: {a.elapsed_time(b) / 5:8.3f} ms")
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
