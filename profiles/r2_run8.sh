mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "rc=$?"; tail -2 gpurun_out/r2_bench_n$N.err | cut -c1-200
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['d2h_gb_per_s_all_gpus'])
print('c3_strong', json.dumps(d.get('c3_strong'))[:700])
print('c5', json.dumps(d.get('c5'))[:600])
PY
