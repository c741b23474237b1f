# A/B of an experiment hook (an environment variable the library reads) on ONE box
# usage: bash profiles/env_ab.sh VAR a b [bench flags]
VAR=$1; A=$2; B=$3; shift 3
for round in 1 2; do
  for v in $A $B; do
    env $VAR=$v timeout 200 python bench.py --no-configs --no-c5 --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', '$*', 'us/step', round(d['ms_per_step']*1e3,2), 'best', round(d['method']['best_ms_per_step']*1e3,2))"
  done
done
