"""Where do a kernel's stall samples / executed instructions sit?  Aggregates the SASS source page of an
.ncu-rep in blocks of N instructions (development tool).
usage: python tools/ncu_sass_regions.py report.ncu-rep [launch_index] [block]"""
import csv, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else -1; B = int(sys.argv[3]) if len(sys.argv) > 3 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; sections.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
sec = sections[which]
hdr, data = sec["hdr"], sec["data"]
isrc, isam, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isam]) for r in data); totx = sum(int(r[iex]) for r in data)
print(sec["name"][:100]); print("samples", tot, "warp-instr", totx, "sass", len(data))
for b in range(0, len(data), B):
    blk = data[b:b + B]
    s = sum(int(r[isam]) for r in blk); x = sum(int(r[iex]) for r in blk)
    ops, st = {}, {}
    for r in blk:
        t = r[isrc].split(); op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
        for i in stall:
            st[hdr[i]] = st.get(hdr[i], 0) + int(r[i] or 0)
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
    tst = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{b:5d} {100*s/max(tot,1):5.1f}% smp {100*x/max(totx,1):5.1f}% ins", top, tst)
