# A/B of two builds of the library on ONE box: glow-tts-train_b200/libmas_ab_{a,b}.so (git-ignored)
# usage: bash profiles/ab_run.sh a b [bench flags]
mkdir -p gpurun_out
A=$1; B=$2; shift 2
cp glow-tts-train_b200/libmas_b200.so /tmp/libmas_keep.so
for round in 1 2 3; do
  for v in $A $B; do
    cp glow-tts-train_b200/libmas_ab_$v.so glow-tts-train_b200/libmas_b200.so
    timeout 200 python bench.py --no-configs --no-c5 --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'ms/step', round(d['ms_per_step']*1e3,2), 'best', round(d['method']['best_ms_per_step']*1e3,2))"
  done
done
cp /tmp/libmas_keep.so glow-tts-train_b200/libmas_b200.so
