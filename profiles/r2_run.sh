mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_gpu.py -x -q 2>&1 | tail -2
for shp in "32 400 2000" "64 400 2000" "32 200 1000" "128 400 2000"; do
  echo "== $shp (by estimate)"; MAS_B200_DEBUG=1 timeout 120 python profiles/time_fused.py $shp 2>&1 | grep -m2 "estimates\|not taken" | cut -c1-150; timeout 120 python profiles/time_fused.py $shp 2>&1 | tail -2
done
