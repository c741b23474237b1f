mkdir -p gpurun_out
out=gpurun_out/r2d_matrix.log; : > $out
for cfg in "15 1" "15 3" "15 5" "12 1" "12 2" "12 3" "12 4"; do
  set -- $cfg
  echo "=== FFMA=$1 TEAMS=$2" >> $out
  MAS_B200_FUSED_FFMA=$1 MAS_B200_FUSED_TEAMS=$2 timeout 120 python profiles/fused_timeline.py 2>&1 | tail -5 >> $out
  MAS_B200_FUSED_FFMA=$1 MAS_B200_FUSED_TEAMS=$2 timeout 120 python profiles/time_fused.py 2>&1 | grep "single" >> $out
  MAS_B200_FUSED_FFMA=$1 MAS_B200_FUSED_TEAMS=$2 timeout 120 python profiles/fused_timeline.py --mean-only 2>&1 | tail -4 >> $out
done
cat $out
