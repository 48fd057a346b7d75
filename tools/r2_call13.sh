#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_stagger.txt; : > $O
for cfg in "MC_TM_STAGGER=0" "MC_TM_STAGGER=1" "MC_TM_STAGGER=0" "MC_TM_STAGGER=1"; do
  echo "== $cfg" >> $O
  env $cfg timeout 120 python tools/tokenmix_bench.py --only fwd 2>&1 | tail -2 >> $O
done
MC_TM_STAGGER=1 timeout 200 python -m pytest tests/test_tokenmix_gpu.py -q 2>&1 | tail -2 >> $O
STEPS=20 bash tools/env_sweep.sh "MC_TM_STAGGER=0" "MC_TM_STAGGER=1" "MC_TM_STAGGER=0" "MC_TM_STAGGER=1" >> $O 2>&1
cat $O
