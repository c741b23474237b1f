mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_model_dropin_gpu.py -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -25 gpurun_out/r2l_pytest.log
timeout 300 python profiles/c5_train_step.py --steps 4 2>&1 | tail -3
