#!/bin/bash
# Sweep of MC_SM_SPLIT (SMs given to the image / text tower while they run on two streams), 1 GPU.
cd "$(dirname "$0")/.."
for split in ${SPLITS:-"" "86,62" "88,60" "84,64" "92,56" "80,68"}; do
  out=$(MC_SM_SPLIT=$split python bench.py --steps 15 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$out" | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('MC_SM_SPLIT=\"$split\" ->', round(d['value']), 'samples/s', round(d['ms_per_step'], 3), 'ms')" || echo "split $split failed"
done
