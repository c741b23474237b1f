mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
MAS_B200_DEBUG=1 timeout 120 python profiles/fused_timeline.py > gpurun_out/r2c_timeline.log 2>&1; tail -8 gpurun_out/r2c_timeline.log
timeout 120 python profiles/time_fused.py > gpurun_out/r2c_time.log 2>&1
timeout 120 python profiles/time_fused.py --ragged >> gpurun_out/r2c_time.log 2>&1
timeout 120 python profiles/time_fused.py 256 400 2000 >> gpurun_out/r2c_time.log 2>&1
tail -12 gpurun_out/r2c_time.log
