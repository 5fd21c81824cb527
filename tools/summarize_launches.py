"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total ns, share.
python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<round>_launches_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
tot = sum(t for _, t, _, _ in rows)
agg = defaultdict(lambda: [0, 0.0])
for name, t, _, _ in rows:
    short = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    short = re.sub(r"^void ", "", short.split("(")[0])[:100]
    agg[short][0] += 1
    agg[short][1] += t
mine = sum(v[1] for k, v in agg.items() if "gvit" in k)
print(f"# {len(rows)} launches, {tot / 1e6:.3f} ms summed kernel time (ncu: cold-cache, serialised - compare SHARES)")
print(f"# libgvit kernels: {mine / 1e6:.3f} ms = {100 * mine / tot:.1f}% of the step")
print(f"{'share':>6} {'ms':>9} {'count':>6}  kernel")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * t / tot:5.1f}% {t / 1e6:9.3f} {c:6d}  {k}")
