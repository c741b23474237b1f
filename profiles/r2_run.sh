mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "rc=$?"; tail -3 gpurun_out/r2m_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')})
print('method', d['method']['replays_ms'], d['method']['best_ms_per_step'])
print('roofline', {k:d['roofline'][k] for k in ('bound','frac','achieved','unit')}, d['roofline']['hbm']['frac'], d['roofline']['fp32_fma']['frac'], d['roofline']['kernel1_alone']['frac'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for k,v in d.get('configs',{}).items():
    if isinstance(v,dict): print(k, v['fused']['ms_per_step'], v['fused']['roofline']['frac'], v['kernel1']['ms_per_step'], v['kernel1']['roofline']['frac'])
print('c3_strong', json.dumps(d.get('c3_strong'))[:600])
print('c5', json.dumps(d.get('c5'))[:400])
print('cpu', json.dumps(d.get('cpu_baseline'))[:1500])
PY
timeout 200 python bench.py --impl reference --steps 10 --warmup 3 | cut -c1-600
