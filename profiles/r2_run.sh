mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -3 gpurun_out/r2_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --mean-only --no-configs --no-c5 > gpurun_out/r2_bench_mean_only.json 2>/dev/null; echo "rc=$?"
timeout 200 python bench.py --workload c1 --no-configs --no-c5 > gpurun_out/r2_bench_c1.json 2>/dev/null; echo "rc=$?"
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference.json 2>/dev/null; echo "rc=$?"
timeout 100 python profiles/fused_timeline.py > gpurun_out/r2_fused_timeline.txt 2>&1
timeout 100 python profiles/fused_timeline.py --mean-only >> gpurun_out/r2_fused_timeline.txt 2>&1
timeout 100 python profiles/fused_timeline.py --random >> gpurun_out/r2_fused_timeline.txt 2>&1
python - <<'PY'
import json
for f in ('r2_bench','r2_bench_mean_only','r2_bench_c1','r2_bench_reference'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'))
PY
