mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; echo "rc=$?"; tail -3 gpurun_out/r2m_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['d2h_gb_per_s_all_gpus'])
print('c3_strong', json.dumps(d.get('c3_strong'))[:900])
print('c5', json.dumps(d.get('c5'))[:500])
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 | cut -c1-300
