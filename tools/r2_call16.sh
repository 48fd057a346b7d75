#!/bin/bash
# round 2, GPU call 16: LayerNorm prologue, vectorised production two tiles ahead - kernel tests + A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_fuse_ln2.txt; : > $O
timeout 300 python -m pytest tests/test_tokenmix_gpu.py -q 2>&1 | tail -6 >> $O
timeout 600 python -m pytest tests/test_model_parity_gpu.py -q 2>&1 | tail -4 >> $O
STEPS=20 bash tools/env_sweep.sh "MC_TM_FUSE_LN=0" "MC_TM_FUSE_LN=1" "MC_TM_FUSE_LN=0" "MC_TM_FUSE_LN=1" >> $O 2>&1
cat $O
