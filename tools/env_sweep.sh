#!/bin/bash
# A/B sweep of environment knobs over the 1-GPU bench.  usage: env_sweep.sh "VAR=1 VAR2=x" "VAR=0" ...  (each argument = one run)
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  out=$(env $cfg python bench.py --steps ${STEPS:-15} --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1)
  echo "$out" | python -c "import sys, json; d = json.loads(sys.stdin.read()); r = d['roofline']; print('$cfg ->', round(d['value']), 'samples/s', round(d['ms_per_step'], 3), 'ms; gemm', round(r['gemm_ms_per_step'], 3), 'ms', round(r['achieved']), 'TF; tokenmix', round(r['token_mix']['ms_per_step'], 3) if r.get('token_mix') else None)" || echo "$cfg -> failed"
done
