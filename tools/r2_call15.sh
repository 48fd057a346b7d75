#!/bin/bash
# round 2, GPU call 15: LayerNorm in the token-mixing prologue - kernel tests, model parity, A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_fuse_ln.txt; : > $O
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -k "row_statistics or streamed" 2>&1 | tail -3 >> $O
timeout 300 python -m pytest tests/test_tokenmix_gpu.py -q 2>&1 | tail -6 >> $O
timeout 600 python -m pytest tests/test_model_parity_gpu.py tests/test_train_step_gpu.py -q 2>&1 | tail -4 >> $O
timeout 600 python -m pytest tests/test_bench_path_gpu.py -q -s -k "benchmarked" 2>&1 | grep -E "^\[|passed|failed" | cut -c1-300 >> $O
STEPS=20 bash tools/env_sweep.sh "MC_TM_FUSE_LN=0" "MC_TM_FUSE_LN=1" "MC_TM_FUSE_LN=0" "MC_TM_FUSE_LN=1" >> $O 2>&1
cat $O
