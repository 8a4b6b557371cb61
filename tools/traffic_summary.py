#!/usr/bin/env python3
"""gpurun_out/traffic_launches.csv (tools/traffic_capture.sh) -> profiles/r2_step_kernel_traffic.json"""
import csv, json, sys
src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/traffic_launches.csv"
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/r2_step_kernel_traffic.json"
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
h = rows[0]
k = {n: h.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
per = {}
for r in rows[1:]:
    v = float(r[k["Metric Value"]].replace(",", ""))
    u = r[k["Metric Unit"]]
    v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1, "ns": 1e-3, "ms": 1e3}.get(u, 1)
    per.setdefault(r[k["ID"]], {"kernel": r[k["Kernel Name"]]})[r[k["Metric Name"]]] = v
ls = list(per.values())
n = len(ls)
rd = sum(l["dram__bytes_read.sum"] for l in ls) / n
wr = sum(l["dram__bytes_write.sum"] for l in ls) / n
us = sum(l["gpu__time_duration.sum"] for l in ls) / n
algo = 46 * (1 << 21)
out = {"kernel": ls[0]["kernel"], "launches": n, "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
       "dram_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": algo, "traffic_over_algorithmic": (rd + wr) / algo,
       "ncu_us_per_launch_serialised": us,
       "how": f"ncu --cache-control none, one pass of dram__bytes_read.sum + dram__bytes_write.sum over {n} consecutive "
              "launches, 4 env groups of 2,097,152 envs round-robin (working set > L2), mean per launch: write-backs of "
              "one launch's dirty lines are counted in the launches that evict them (tools/traffic_capture.sh)"}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
