#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_early.txt; : > $O
for cfg in "MC_TM_EARLY=0 MC_TM_STAGGER=0" "MC_TM_EARLY=1 MC_TM_STAGGER=0" "MC_TM_EARLY=0 MC_TM_STAGGER=1" "MC_TM_EARLY=1 MC_TM_STAGGER=1" "MC_TM_EARLY=0 MC_TM_STAGGER=0" "MC_TM_EARLY=1 MC_TM_STAGGER=0"; do
  echo "== $cfg" >> $O
  env $cfg timeout 120 python tools/tokenmix_bench.py --only fwd 2>&1 | tail -2 >> $O
done
MC_TM_EARLY=1 timeout 200 python -m pytest tests/test_tokenmix_gpu.py -q 2>&1 | tail -2 >> $O
MC_TM_EARLY=1 MC_TM_STAGGER=1 timeout 200 python -m pytest tests/test_tokenmix_gpu.py -q 2>&1 | tail -2 >> $O
cat $O
