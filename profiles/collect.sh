#!/bin/bash
# Round profile collection (run under gpurun, one GPU).  Usage: bash profiles/collect.sh rN
# 1. plain bench run (must exit 0), 2. launch list of the SAME command (gpu__time_duration),
# 3. one `--set full` capture of kernel (1) and one of the logp kernel.  Reports land in gpurun_out/.
set -u
R=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${R}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:systolic -s 6 -c 1 -f -o gpurun_out/${R}_prof_mas python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_mas.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused -s 6 -c 1 -f -o gpurun_out/${R}_prof_fused $CMD > gpurun_out/${R}_ncu_logp.log 2>&1
echo collected
