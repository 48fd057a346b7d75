#!/bin/bash
# round 2, GPU call 6: which kernel family is wrong when its grid is sized for a share of the SMs (MC_SM_LIMIT)?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_smlimit_tests.txt; : > $O
for lim in 84 64 80 68 48 100; do
  for f in tests/test_gemm_gpu.py tests/test_tokenmix_gpu.py tests/test_rowwise_gpu.py; do
    echo "== MC_SM_LIMIT=$lim $f" >> $O
    MC_SM_LIMIT=$lim timeout 300 python -m pytest $f -q 2>&1 | grep -E "passed|failed|FAILED" | head -12 >> $O
  done
done
echo "== MC_SM_LIMIT=84 model parity B32" >> $O
MC_SM_LIMIT=84 timeout 600 python -m pytest tests/test_model_parity_gpu.py -q -k "b32_full or S2" 2>&1 | grep -E "passed|failed|FAILED" | head >> $O
MC_SM_LIMIT=64 timeout 600 python -m pytest tests/test_model_parity_gpu.py -q -k "b32_full or S2" 2>&1 | grep -E "passed|failed|FAILED" | head >> $O
cat $O
