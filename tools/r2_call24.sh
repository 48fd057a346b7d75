#!/bin/bash
# round 2, GPU call 24: LayerNorm prologue on/off with the spill-free store pass (interleaved A/B), and the GPU suite with it off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_call24.txt; : > $O
STEPS=20 bash tools/env_sweep.sh "MC_TM_FUSE_LN=0" "MC_TM_FUSE_LN=1" "MC_TM_FUSE_LN=0" "MC_TM_FUSE_LN=1" >> $O 2>&1
MC_TM_FUSE_LN=0 timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -2 >> $O
cat $O
