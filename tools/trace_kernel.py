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
buf = torch.zeros(16 * per_warp + 2048, dtype=torch.int64, device=dev)   # + begin / end of up to 1024 CTAs
# warm up without tracing, then trace exactly one launch
bench.kernel_rooflines(dev, args.batch, bench.load_peaks(), iters=2, only=[args.name])
assert fn(buf.data_ptr(), per_warp) == 0
res = bench.kernel_rooflines(dev, args.batch, bench.load_peaks(), iters=1, only=[args.name])
torch.cuda.synchronize()
assert fn(None, 0) == 0
raw_all = buf.cpu().numpy().astype("uint64")
raw, span = raw_all[: 16 * per_warp], raw_all[16 * per_warp:].reshape(-1, 2)
ev = sorted(((int(v) & 0xFFFFFFFFFFF, (int(v) >> 44) >> 8, (int(v) >> 44) & 0xff) for v in raw if int(v) != 0))
n = len(ev)
print(f"# {args.name}: {res[args.name]['ms']*1e3:.1f} us/launch, {n} events from CTA 0 (showing launches after warm-up; clock cycles)")
# the setter was armed before 3 warm-up + 1 timed launches of kernel_rooflines: split launches by large gaps
t0 = ev[0][0] if ev else 0
last = t0
for t, warp, eid in ev[: args.max]:
    print(f"{t - t0:10d} (+{t - last:6d})  warp {warp:2d}  ev {eid}")
    last = t

# per-CTA spans (globaltimer ns) of the LAST traced launch: launch skew, imbalance, tail
live = [(int(b), int(e)) for b, e in span if b != 0 and e != 0]
if live:
    t00 = min(b for b, _ in live)
    late = sorted(((int(e) - t00, i) for i, (b, e) in enumerate(span) if b != 0 and e != 0), reverse=True)[:12]
    print("# latest CTAs (end ns, blockIdx.x): " + ", ".join(f"{t}:{i}" for t, i in late))
    t0 = min(b for b, _ in live)
    ends = sorted(e - t0 for _, e in live)
    begins = sorted(b - t0 for b, _ in live)
    durs = sorted(e - b for b, e in live)
    q = lambda v, f: v[min(len(v) - 1, int(f * len(v)))]
    print(f"# spans of {len(live)} CTAs (ns from the first CTA's begin): begin median {q(begins, .5)} max {begins[-1]}; "
          f"end min {ends[0]} p25 {q(ends, .25)} median {q(ends, .5)} p75 {q(ends, .75)} max {ends[-1]}; "
          f"busy min {durs[0]} median {q(durs, .5)} max {durs[-1]}")
