import json, sys
t = json.load(open(sys.argv[1]))
tot = sum(v["ms_total_per_step"] for v in t.values())
print("total eager-timed ms/step", round(tot, 3))
for k, v in t.items():
    print(f"{v['ms_total_per_step']:8.3f} ms  x{v['launches_per_step']:3d}  {v['us_avg']:8.1f} us  {k}")
