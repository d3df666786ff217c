"""Top SASS instructions by stall samples with context (development tool).
usage: python tools/ncu_top_sass.py report.ncu-rep [n_top] [context]"""
import csv, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25; ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 2
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
    elif hdr is not None and len(r) == len(hdr):
        data.append(r)
isrc, isam, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isam]) for r in data)
order = sorted(range(len(data)), key=lambda i: -int(data[i][isam]))[:ntop]
for i in sorted(order):
    print("----")
    for j in range(max(0, i - ctx), min(len(data), i + ctx + 1)):
        r = data[j]
        st = sorted(((int(r[k] or 0), hdr[k][6:]) for k in stall), reverse=True)[:2]
        mark = ">>" if j == i else "  "
        print(f"{mark}{j:5d} {100*int(r[isam])/tot:5.2f}% x{int(r[iex]):>9}  {r[isrc].strip()[:90]:90s} {st}")
