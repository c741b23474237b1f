mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_gpu.py -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -4 gpurun_out/r2k_pytest.log
timeout 100 python profiles/fused_timeline.py 2>&1 | tail -9
timeout 100 python profiles/time_fused.py 2>&1 | tail -2
