"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (for profiles/).
usage: python tools/launch_summary.py launches.csv "<header comment>" > profiles/<name>.txt"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); u = r[iu]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    name = r[ik].split("(")[0][:90]
    a = agg.setdefault(name, [0.0, 0]); a[0] += v; a[1] += 1
tot = sum(a[0] for a in agg.values()); n = sum(a[1] for a in agg.values())
for line in sys.argv[2:]:
    print("# " + line)
print(f"# total {tot:.1f} us over {n} launches (cold-cache, serialised: compare shares)")
for name, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{us:10.1f} us  {100*us/tot:5.1f}%  x{c:4d}  {name}")
