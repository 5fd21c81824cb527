"""Event trace of CTA 0 of a warp-specialised libgvit kernel (needs the -DGVIT_TRACE build:
   make EXTRA=-DGVIT_TRACE OBJ_DIR=build/obj_trace LIB=tools/bin/libgvit_trace.so).
python tools/trace_kernel.py attn_bwd [--batch 256] > gpurun_out/trace_attn_bwd.txt"""
import argparse, ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_augmented_vision_transformers_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libgvit_trace.so")   # never next to the product library
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("name")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--max", type=int, default=400)
args = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _lib.load()
setter = {"attn_bwd": "gvit_debug_set_trace_attn", "attn_fwd": "gvit_debug_set_trace_attn",
          "agg_fwd": "gvit_debug_set_trace_agg" if os.environ.get("GVIT_AGG_NOPAIR") else "gvit_debug_set_trace_agg4", "knn_fwd": "gvit_debug_set_trace_knn", "graph_bwd": "gvit_debug_set_trace_graph_bwd_pair"}[args.name]
fn = getattr(lib, setter); fn.argtypes = [ctypes.c_void_p, ctypes.c_uint]; fn.restype = ctypes.c_int
per_warp = 4096
buf = torch.zeros(16 * per_warp, dtype=torch.int64, device=dev)
# warm up without tracing, then trace exactly one launch
bench.kernel_rooflines(dev, args.batch, bench.load_peaks(), iters=2, only=[args.name])
assert fn(buf.data_ptr(), per_warp) == 0
res = bench.kernel_rooflines(dev, args.batch, bench.load_peaks(), iters=1, only=[args.name])
torch.cuda.synchronize()
assert fn(None, 0) == 0
raw = buf.cpu().numpy().astype("uint64")
ev = sorted(((int(v) & 0xFFFFFFFFFFF, (int(v) >> 44) >> 8, (int(v) >> 44) & 0xff) for v in raw if int(v) != 0))
n = len(ev)
print(f"# {args.name}: {res[args.name]['ms']*1e3:.1f} us/launch, {n} events from CTA 0 (showing launches after warm-up; clock cycles)")
# the setter was armed before 3 warm-up + 1 timed launches of kernel_rooflines: split launches by large gaps
t0 = ev[0][0] if ev else 0
last = t0
for t, warp, eid in ev[: args.max]:
    print(f"{t - t0:10d} (+{t - last:6d})  warp {warp:2d}  ev {eid}")
    last = t
