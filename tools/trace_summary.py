"""Per-warp view of a tools/trace_kernel.py trace: python tools/trace_summary.py <trace.txt> <warp> [<warp> ...]"""
import re
import sys

ev = []
for l in open(sys.argv[1]):
    m = re.match(r"\s*(\d+) \(\+\s*(-?\d+)\)\s+warp\s+(\d+)\s+ev (\d+)", l)
    if m:
        ev.append((int(m.group(1)), int(m.group(3)), int(m.group(4))))
for w in map(int, sys.argv[2:]):
    print("--- warp", w)
    last = None
    for t, ww, e in ev:
        if ww == w:
            print(f"{t:8d} +{(t - last) if last is not None else 0:6d} ev{e}")
            last = t
