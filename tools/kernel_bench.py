"""Launch one libgvit kernel a few times at the bench shape (B=256, ViT-B/16, k=8) - the short command line that
`ncu --set full -k regex:<kernel>` profiles.  python tools/kernel_bench.py <name> [--iters 3] [--batch 256]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("names", nargs="*", default=[])
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--batch", type=int, default=256)
args = ap.parse_args()
dev = torch.device("cuda", 0)
res = bench.kernel_rooflines(dev, args.batch, bench.load_peaks(), iters=args.iters, only=args.names or None)
for n, d in res.items():
    print(json.dumps({"kernel": n, **{k: (round(v, 5) if isinstance(v, float) else v) for k, v in d.items()}}))
