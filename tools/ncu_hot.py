"""Print the hottest SASS instructions of an .ncu-rep with their neighbourhood and the dominant stall reason.
python tools/ncu_hot.py rep.ncu-rep [top_n] [context]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 8; ctxn = int(sys.argv[3]) if len(sys.argv) > 3 else 5
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; body = []
for r in rows:
    if r and r[0] == "Address": hdr = r
    elif hdr and r and r[0].startswith("0x"): body.append(r)
si, so, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in body) or 1
order = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:topn]
for i in order:
    r = body[i]
    st = sorted(((int(r[c]), hdr[c]) for c in stall_cols), reverse=True)[:3]
    print(f"--- {100*int(r[si])/tot:.1f}% of samples, top stalls: " + ", ".join(f"{n}={v}" for v, n in st if v))
    for j in range(max(0, i - ctxn), min(len(body), i + 3)):
        mark = ">>" if j == i else "  "
        print(f" {mark} {body[j][0][-5:]} s={body[j][si]:>5s} x={body[j][ie]:>8s} {body[j][so].strip()[:110]}")
