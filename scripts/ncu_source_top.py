"""Summarise an `ncu --page source --csv` export: per kernel, total stall reasons and the hottest SASS lines.
usage: python scripts/ncu_source_top.py gpurun_out/prof_X_source.csv.gz [kernel-substring] [top-n]"""
import csv, gzip, sys, collections

path = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
op = gzip.open if path.endswith(".gz") else open
kernels = []  # (name, header, rows)
with op(path, "rt", newline="") as f:
    cur = None
    for row in csv.reader(f):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = [row[1], None, []]
            kernels.append(cur)
        elif row[0] == "Address":
            cur[1] = row
        elif cur is not None and cur[1] is not None:
            cur[2].append(row)
seen = collections.Counter()
for name, hdr, rows in kernels:
    if filt not in name:
        continue
    seen[name] += 1
    if seen[name] > 1:
        continue
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    nsamp = 0
    for r in rows:
        nsamp += int(r[idx["# Samples"]] or 0)
        for c in stall_cols:
            tot[c] += int(r[idx[c]] or 0)
    print(f"=== {name[:90]}  samples={nsamp} sass_lines={len(rows)}")
    print("  stalls: " + ", ".join(f"{k[6:]}={v * 100 // max(1, nsamp)}%" for k, v in tot.most_common(8)))
    rows_s = sorted(enumerate(rows), key=lambda ir: -int(ir[1][idx["# Samples"]] or 0))[:topn]
    for i, r in sorted(rows_s):
        s = int(r[idx["# Samples"]] or 0)
        top = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"  {i:5d} {s * 100.0 / max(1, nsamp):5.1f}%  {r[idx['Source']].strip()[:70]:70s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
