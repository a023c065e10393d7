"""summarise an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum) of scripts/gpu_solve_prof.py: last solve only"""
import csv, collections, sys, re
path = sys.argv[1]; nsolves = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lines = [l for l in open(path) if not l.startswith("==")]
by = {}
for row in csv.DictReader(lines):
    i = int(row["ID"]); d = by.setdefault(i, {"name": row["Kernel Name"].split("(")[0].replace("void ", ""), "grid": int(re.findall(r"\d+", row["Grid Size"])[0])})
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    if "time" in row["Metric Name"]:
        d["us"] = v / 1e3 if u.startswith("ns") else v if u.startswith("us") else v * 1e3
    else:
        d["mb"] = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u] * v
ids = sorted(by); ids = ids[len(ids) * (nsolves - 1) // nsolves:]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for i in ids:
    d = by[i]; a = agg[d["name"]]; a[0] += 1; a[1] += d["us"]; a[2] += d.get("mb", 0)
print(f"{'kernel':40s} {'n':>5s} {'us':>9s} {'MB(dram rd)':>12s} {'GB/s':>8s}")
for k, (n, us, mb) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:40s} {n:5d} {us:9.1f} {mb:12.1f} {mb / us * 1e3:8.1f}")
print(f"total kernel time {sum(a[1] for a in agg.values()):.1f} us, dram read {sum(a[2] for a in agg.values()):.1f} MB")
for kn in ("solve_fwd_step_kernel", "solve_bwd_step_kernel"):
    for lo, hi in ((0, 400), (400, 3000), (3000, 10 ** 9)):
        sel = [by[i] for i in ids if kn in by[i]["name"] and lo <= by[i]["grid"] < hi]
        if sel:
            us = sorted(d["us"] for d in sel); mb = sum(d.get("mb", 0) for d in sel)
            print(f"{kn} grid[{lo},{hi}): n={len(sel)} total={sum(us):.1f}us median={us[len(us) // 2]:.1f}us min={us[0]:.1f} max={us[-1]:.1f} MB={mb:.0f} GB/s={mb / sum(us) * 1e3:.0f}")
