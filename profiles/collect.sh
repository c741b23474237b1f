#!/bin/bash
# Round profile collection (run under gpurun, one GPU, ONE profiler pass per call).
#   bash profiles/collect.sh rN list    plain bench run (must exit 0), then the launch list of the SAME command
#   bash profiles/collect.sh rN mas     plain run, then one `--set full` capture of kernel (1) at C1
#   bash profiles/collect.sh rN fused   ... of the single-launch kernel (2) at C2
#   bash profiles/collect.sh rN logp    ... of the log-likelihood kernel alone at C2
# Reports land in gpurun_out/; `python profiles/summarize.py rN` turns them into the tracked text files.
set -u
R=${1:-r1}
WHAT=${2:-list}
mkdir -p gpurun_out
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-c5"
FULL="ncu --set full --clock-control none --import-source on -c 1 -f"
case "$WHAT" in
  list)
    $BENCH > gpurun_out/${R}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain.log; exit 1; }
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $BENCH > gpurun_out/${R}_ncu_list.log 2>&1 ;;
  mas)
    $BENCH --workload c1 > gpurun_out/${R}_plain_c1.log 2>&1 || { echo "plain run failed"; exit 1; }
    $FULL -k regex:systolic -s 6 -o gpurun_out/${R}_prof_mas $BENCH --workload c1 > gpurun_out/${R}_ncu_mas.log 2>&1 ;;
  fused)
    $BENCH > gpurun_out/${R}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
    $FULL -k regex:fused -s 6 -o gpurun_out/${R}_prof_fused $BENCH > gpurun_out/${R}_ncu_fused.log 2>&1 ;;
  logp)
    python profiles/run_logp.py > gpurun_out/${R}_plain_logp.log 2>&1 || { echo "plain logp run failed"; exit 1; }
    $FULL -k regex:mas_logp -s 8 -o gpurun_out/${R}_prof_logp python profiles/run_logp.py > gpurun_out/${R}_ncu_logp_alone.log 2>&1 ;;
  *) echo "unknown stage $WHAT"; exit 2 ;;
esac
echo "collected $WHAT"
