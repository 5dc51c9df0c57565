"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  usage: python scripts/ncu_launch_summary.py file.csv [top]"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
idx = {k: j for j, k in enumerate(rows[h])}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) < len(rows[h]) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[idx["Metric Value"]].replace(",", ""))
    u = r[idx["Metric Unit"]]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
    grid = r[idx["Grid Size"]] if "Grid Size" in idx else ""
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f}%  x{v[0]:4d}  {v[1] / v[0]:8.1f}  {k[:90]}")
