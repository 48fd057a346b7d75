#!/bin/bash
# round 2, GPU call 3: whole GPU suite (no -x), token-mix segment-width sweep, PDL A/B, text-tower tile-width sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | grep -E "^\[|passed|failed|FAILED|Error|EXEMPT|skipped" | tail -80 > gpurun_out/r2_gpu_tests2.log
T=gpurun_out/r2_tokenmix_sweep.txt; : > $T
for cfg in "MC_TM_NO_AUG=0" "MC_TM_NO_AUG=1" "MC_TM_SW_FWD=64" "MC_TM_SW_DGRAD=64" "MC_TM_SW_DGRAD=128" "MC_TM_SW_WGRAD=64" "MC_TM_SW_FWD=64 MC_TM_SW_DGRAD=64 MC_TM_SW_WGRAD=64"; do
  echo "== $cfg" >> $T
  env $cfg timeout 120 python tools/tokenmix_bench.py 2>&1 | tail -6 >> $T
done
env MC_TM_SW_FWD=64 MC_TM_SW_DGRAD=64 MC_TM_SW_WGRAD=64 timeout 200 python -m pytest tests/test_tokenmix_gpu.py -q 2>&1 | tail -2 >> $T
G=gpurun_out/r2_gemm_bn.txt; : > $G
for bn in 0 128 192; do
  echo "== MC_GEMM_BN=$bn (0 = cost model)" >> $G
  env MC_GEMM_BN=$bn timeout 120 python tools/gemm_bench.py txt_lin3 txt_lin4 txt_dz2 txt_dv txt_dw3 txt_dw4 lin3 lin4 dz2 dv >> $G 2>&1
done
S=gpurun_out/r2_pdl_ab.txt; : > $S
MC_PDL=1 timeout 300 python -m pytest tests/test_train_step_gpu.py tests/test_gemm_gpu.py tests/test_tokenmix_gpu.py -x -q 2>&1 | tail -2 >> $S
STEPS=15 bash tools/env_sweep.sh "MC_PDL=0" "MC_PDL=1" "MC_PDL=0" "MC_PDL=1" >> $S 2>&1
tail -12 gpurun_out/r2_gpu_tests2.log; cat $S
