mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_gpu.py -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -2 gpurun_out/r2p_pytest.log
timeout 100 python profiles/fused_timeline.py 2>&1 | tail -12 | grep -v "backtrack thread"
timeout 100 python profiles/time_fused.py 2>&1 | tail -2
MAS_B200_FUSED_MODE=cluster timeout 100 python profiles/time_fused.py 32 400 2000 2>&1 | tail -2
timeout 200 python bench.py --no-configs --no-c5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench ms/step', d['ms_per_step'], d['value'])"
timeout 200 python bench.py --mean-only --no-configs --no-c5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench mean_only ms/step', d['ms_per_step'], d['value'])"
