"""TMA streaming bandwidth for the tile shapes the kernels use (diagnostic).  python tools/tma_probe.py"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_augmented_vision_transformers_b200 import _lib
lib = ctypes.CDLL(_lib.LIB_PATH)
f = lib.gvit_probe_tma
f.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_void_p]
dev = torch.device("cuda", 0)
sink = torch.zeros(1, dtype=torch.int64, device=dev)
for (B, rows, cols) in [(256, 197, 2304), (256, 196, 768)]:
    xs = [torch.randn(B, rows, cols, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    nbytes = xs[0].numel() * 2
    for box in (64, 128, 208, 256):
        for stages in (2, 4, 6):
            for cps in (1, 2):
                if stages * box * 128 * cps > 200 * 1024:
                    continue
                st = torch.cuda.current_stream().cuda_stream
                for i in range(2):
                    rc = f(xs[i % 2].data_ptr(), B, rows, cols, box, stages, cps, sink.data_ptr(), st)
                    assert rc == 0, rc
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for i in range(4):
                    f(xs[i % 2].data_ptr(), B, rows, cols, box, stages, cps, sink.data_ptr(), st)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 4
                # bytes actually moved: boxes cover ceil(rows/box)*box rows, OOB rows cost no DRAM traffic
                print(f"tensor ({B},{rows},{cols}) box_rows={box:3d} stages={stages} ctas/sm={cps}: {ms*1e3:8.1f} us  {nbytes/ms/1e6:7.0f} GB/s")
