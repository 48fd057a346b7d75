#!/bin/bash
# final defaults: GPU suite + default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.txt 2>&1; tail -1 gpurun_out/r2_final_smoke.txt
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | grep -E "^\[|passed|failed|FAILED|Error|EXEMPT|skipped|losses" | cut -c1-400 > gpurun_out/r2_final_gpu_tests.log; tail -2 gpurun_out/r2_final_gpu_tests.log
timeout 400 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 400 gpurun_out/r2_final_bench.json
