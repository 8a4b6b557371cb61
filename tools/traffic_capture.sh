#!/bin/bash
# DRAM bytes per launch of the headline kernel in its benchmark setting: 4 env groups of 2 Mi envs stepped round-robin
# (working set > L2), caches left alone between kernels (--cache-control none), one metric pass (no replay), so the
# write-back of one launch's dirty lines shows up during the following launches.  Writes
# gpurun_out/traffic_launches.csv; tools/traffic_summary.py turns it into profiles/r2_step_kernel_traffic.json.
set -e
cd "$(dirname "$0")/.."
python tools/profile_step.py --groups 4 --steps 64 > gpurun_out/traffic_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k regex:step_fast -s 1200 -c 64 --csv --log-file gpurun_out/traffic_launches.csv \
    python tools/profile_step.py --groups 4 --steps 64 > gpurun_out/traffic_ncu.log 2>&1
tail -2 gpurun_out/traffic_ncu.log
