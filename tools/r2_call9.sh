#!/bin/bash
# round 2, GPU call 9: rolling L2 prefetch A/B, tests after the interleaved backward enqueue
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_l2pf.txt; : > $O
for d in 0 4 8 16; do
  echo "== MC_GEMM_L2PF=$d" >> $O
  MC_GEMM_L2PF=$d timeout 200 python tools/gemm_bench.py lin3 lin4 dz2 dv dw3 dw4 txt_lin3 txt_lin4 txt_dz2 txt_dv txt_dw3 >> $O 2>&1
done
MC_GEMM_L2PF=8 timeout 300 python -m pytest tests/test_gemm_gpu.py -q 2>&1 | tail -2 >> $O
STEPS=20 bash tools/env_sweep.sh "MC_GEMM_L2PF=0" "MC_GEMM_L2PF=8" "MC_GEMM_L2PF=0" "MC_GEMM_L2PF=8" "MC_GEMM_L2PF=16" >> $O 2>&1
timeout 600 python -m pytest tests/test_train_step_gpu.py tests/test_bench_path_gpu.py -q -s 2>&1 | grep -E "^\[|passed|failed|FAILED" | cut -c1-300 | tail -12 >> $O
cat $O
