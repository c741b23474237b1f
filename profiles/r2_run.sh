mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_gpu.py -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -3 gpurun_out/r2o_pytest.log
timeout 100 python profiles/fused_timeline.py 2>&1 | tail -13 | grep -v "sweep warp"
timeout 100 python profiles/time_fused.py 2>&1 | tail -2
timeout 200 python bench.py --no-configs --no-c5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench ms/step', d['ms_per_step'], d['value'])"
