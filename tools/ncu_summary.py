#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.json + a launch-share table.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_step_full.json
    python tools/ncu_summary.py --launches gpurun_out/launches.csv profiles/r1_launches.md
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                v = r[hdr.index(k)].replace(",", "")
                try:
                    v = float(v)
                except ValueError:
                    pass
                d[k] = {"value": v, "unit": units[hdr.index(k)]}
        res.append(d)
    json.dump({"report": rep, "launches": res}, open(out, "w"), indent=1)
    for d in res:
        t = d["gpu__time_duration.sum"]
        rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        print(d["kernel"][:50], t["value"], t["unit"], "dram rd", rd["value"], rd["unit"], "wr", wr["value"], wr["unit"],
              "regs", d["launch__registers_per_thread"]["value"],
              "inst", d["smsp__inst_executed.sum"]["value"],
              "issue%", d["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"])


def launches(path, out):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        k = r["Kernel Name"]
        tot[k][0] += 1
        tot[k][1] += v
    total = sum(v[1] for v in tot.values())
    with open(out, "w") as f:
        f.write(f"# kernel launch list ({path}; ncu --metrics gpu__time_duration.sum, cold-cache, serialised)\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:90]}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / total:.1f}% |\n")
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[1], sys.argv[2])
