mkdir -p gpurun_out
: > gpurun_out/r2_sweep_k.txt
for shp in "32 200 1000" "148 200 1000" "592 200 1000" "256 400 2000" "32 400 2000" "8 1024 8192" "32 1024 8192"; do
  timeout 120 python profiles/sweep_k.py $shp 0 2>&1 | tail -1 >> gpurun_out/r2_sweep_k.txt
done
cat gpurun_out/r2_sweep_k.txt
