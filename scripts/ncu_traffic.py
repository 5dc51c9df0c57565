"""profiles/traffic.json from an `ncu --set full --page raw --csv` export: per entry point, the DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) of its main kernel per launch (median over the captured launches).
usage: python scripts/ncu_traffic.py gpurun_out/prof_X_raw.csv [profiles/traffic.json]"""
import csv, json, statistics, sys

KERNELS = {"mca_attn_bwd": "attn_bwd_kernel", "mca_attn_fwd": "attn_fwd_kernel", "mca_gemm_bf16": "gemm2_tc_kernel"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
i_name, i_r, i_w, i_t = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
out = {}
for ep, k in KERNELS.items():
    vals, durs = [], []
    for r in rows[2:]:
        if k in r[i_name]:
            vals.append(float(r[i_r]) * UNIT[units[i_r]] + float(r[i_w]) * UNIT[units[i_w]])
            durs.append(float(r[i_t]))
    if vals:
        out[ep] = {"kernel": k, "dram_bytes_per_launch": statistics.median(vals), "launches_captured": len(vals),
                   "ncu_duration_median": statistics.median(durs), "ncu_duration_unit": units[i_t], "source": sys.argv[1]}
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/traffic.json"
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
