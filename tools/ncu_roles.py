"""Per-role stall attribution of gemm_tc_kernel from an ncu source page (roles split at USETMAXREG)."""
import csv, subprocess, sys, collections
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[idx["# Samples"]]) for r in data)
b = [0] + [i for i, r in enumerate(data) if "USETMAXREG" in r[idx["Source"]]] + [len(data)]
names = ["prologue", "mma", "producer", "accumulate+epilogue"]
for (x, y), nm in zip(zip(b[:-1], b[1:]), names):
    n = sum(int(r[idx["# Samples"]]) for r in data[x:y])
    agg = {c: sum(int(r[idx[c]]) for r in data[x:y]) for c in stall}
    ie = sum(int(r[idx["Instructions Executed"]]) for r in data[x:y])
    print(f"{nm:22s} samples {n:7d} ({100*n/tot:4.1f}%) warp-instr {ie:11d}  " + ", ".join(f"{k[6:]} {100*v/max(n,1):.0f}%" for v, k in sorted(((v, k) for k, v in agg.items()), reverse=True)[:4]))
top = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:14]
for i in sorted(top):
    r = data[i]; n = int(r[idx["# Samples"]])
    print(f"  {i:5d} {100*n/tot:5.1f}%  {r[idx['Source']].strip()[:70]}")
