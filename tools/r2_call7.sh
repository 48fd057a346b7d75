#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_gemm_limit_probe.txt; : > $O
for lim in 48 80 84 64 68 148; do
  MC_SM_LIMIT=$lim MC_GEMM_DEBUG=1 timeout 300 python tools/gemm_limit_probe.py >> $O 2>&1
done
grep -E "FAIL|ok" $O | head -120
