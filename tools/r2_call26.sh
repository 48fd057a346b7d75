#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_model_parity_gpu.py -q -x -k "layernorm_in_the_token_mixing_prologue" 2>&1 | tail -25 | cut -c1-300 | tee gpurun_out/r2j_ln_prologue_test.log
