"""Per-source-line executed warp instructions and stall samples from an .ncu-rep (development tool).
usage: python tools/ncu_lines.py report.ncu-rep [min_pct] [kernel_index]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5; which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; agg = collections.OrderedDict(); k = -1
for r in rows:
    if r and r[0] == "Kernel Name":
        k += 1; continue
    if k != which:
        continue
    if r and r[0] == "Line No" and len(r) > 8:
        hdr = r; isam = hdr.index("# Samples"); iex = hdr.index("Instructions Executed"); continue
    if hdr is not None and len(r) == len(hdr) and r[0].isdigit():
        a = agg.setdefault(int(r[0]), [r[1], 0, 0, 0])
        a[1] += int(r[isam] or 0); a[2] += int(r[iex] or 0); a[3] += 1
tot = sum(a[1] for a in agg.values()); totx = sum(a[2] for a in agg.values())
print("samples", tot, "warp-instr", totx)
for ln in sorted(agg):
    src, s, x, n = agg[ln]
    if 100 * s / max(tot, 1) >= minp or 100 * x / max(totx, 1) >= minp:
        print(f"{ln:>5} {100*s/max(tot,1):5.1f}% smp {100*x/max(totx,1):5.1f}% ins {n:4d} sass  {src.strip()[:100]}")
