mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_training_gpu.py tests/test_consumers_gpu.py -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -25 gpurun_out/r2n_pytest.log | grep -v Warning
