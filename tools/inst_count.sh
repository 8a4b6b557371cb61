#!/bin/bash
# Dynamic warp-instruction count and duration of the last step_fast_kernel launches of tools/profile_step.py under ncu
# (a number taken under ncu is never a bench value; this is for comparing builds).
#   tools/inst_count.sh [LIB.so] [ROOT] [extra profile_step args]
lib=${1:-}; root=${2:-.}; shift 2 2>/dev/null
[ -n "$lib" ] && export SHIPENV_LIB=$lib
cd $root
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:step_fast --launch-skip 300 -c 3 --csv python tools/profile_step.py --steps 3 "$@" 2>/dev/null \
  | python -c "
import csv, sys
rows = [r for r in csv.reader(sys.stdin) if len(r) > 10]
hdr = rows[0]; k = {n: hdr.index(n) for n in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID')}
out = {}
for r in rows[1:]:
    out.setdefault(r[k['ID']], {})[r[k['Metric Name']]] = r[k['Metric Value']]
for i, m in out.items():
    print('${lib:-default}', i, ' '.join(f'{a.split(\".\")[0].replace(\"smsp__\",\"\")}={v}' for a, v in m.items()))
"
