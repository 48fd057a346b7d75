#!/bin/bash
# round 2, last GPU call: the shipped defaults (LayerNorm prologue off) - smoke(), the model-parity file incl. the opt-in prologue test, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.txt 2>&1; tail -1 gpurun_out/r2j_smoke.txt
timeout 300 python -m pytest tests/test_model_parity_gpu.py -q -k "b32 or golden_S2" 2>&1 | tail -2 | tee gpurun_out/r2j_parity_tests.log
timeout 300 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 300 gpurun_out/r2j_bench.json
