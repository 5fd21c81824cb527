"""Condense an .ncu-rep (ncu --set full --import-source on) into text: headline metrics per kernel + the hottest SASS
lines by stall samples.  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n] > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
seen = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    seen.setdefault(name, r)          # first captured launch of each kernel
for name, r in seen.items():
    print("=" * 100)
    print("kernel:", name[:160])
    for w in WANT:
        if w in hdr:
            print(f"  {w:85s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
done = set()
for b in blocks:
    if b["name"] in done or "hdr" not in b:
        continue
    done.add(b["name"])
    h = b["hdr"]
    si, so, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    tot = sum(int(r[si]) for r in b["rows"]) or 1
    print("=" * 100)
    print(f"hot SASS of {b['name'][:120]}  ({tot} stall samples, {len(b['rows'])} instructions)")
    for r in sorted(b["rows"], key=lambda r: -int(r[si]))[:topn]:
        print(f"  {100 * int(r[si]) / tot:5.1f}%  exec={r[ie]:>9s}  {r[so].strip()[:100]}")
