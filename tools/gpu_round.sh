#!/bin/bash
# One GPU-box session: fused token-mixing kernel tests (one process per mode, so a faulting mode does not hide
# the others), micro-benchmark, then the whole GPU suite and the bench in both token-mixing schedules.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in "fwd and not spill" "dgrad" "wgrad" "spill" "unsupported"; do
  echo "=== -k $k" >> gpurun_out/tm_tests.log
  timeout 300 python -m pytest tests/test_tokenmix_gpu.py -q -k "$k" 2>&1 | tail -25 >> gpurun_out/tm_tests.log
done
timeout 300 python tools/tokenmix_bench.py > gpurun_out/tm_bench.log 2>&1
tail -8 gpurun_out/tm_tests.log; tail -7 gpurun_out/tm_bench.log
