"""Same-box A/B of the fused-epilogue GEMMs between the product library and an older build:
   git stash; make -j OBJ_DIR=build/obj_old LIB=tools/bin/libgvit_old.so; git stash pop; make -j
   gpurun -- 'OLD=old python tools/ab_proj.py; python tools/ab_proj.py'"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_augmented_vision_transformers_b200 import _lib  # noqa: E402

if os.environ.get("OLD"):
    _lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", os.environ["OLD"] if os.environ["OLD"].endswith(".so") else "libgvit_old.so")
import torch  # noqa: E402

import bench  # noqa: E402

res = bench.kernel_rooflines(torch.device("cuda", 0), 256, bench.load_peaks(), iters=10,
                             only=["proj_fused", "proj_fused_bf16_stream", "fc2_fused", "fc1_fused", "fc2_bwd_fused"])
for n, d in res.items():
    print(os.environ.get("OLD", "new"), n, round(d["ms"], 5))
