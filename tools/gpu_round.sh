#!/bin/bash
# One GPU-box session: fused token-mixing kernel tests (one process per mode, so a faulting mode does not hide
# the others), micro-benchmark, then the whole GPU suite and the bench in both token-mixing schedules.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in "fwd" "dgrad" "wgrad" "unsupported"; do
  echo "=== -k $k" >> gpurun_out/tm_tests.log
  timeout 300 python -m pytest tests/test_tokenmix_gpu.py -q -k "$k" 2>&1 | tail -25 >> gpurun_out/tm_tests.log
done
timeout 300 python tools/tokenmix_bench.py > gpurun_out/tm_bench.log 2>&1
tail -8 gpurun_out/tm_tests.log; tail -7 gpurun_out/tm_bench.log
if [ "$1" = "full" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
  timeout 600 python bench.py > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; tail -c 1500 gpurun_out/bench_fused.json
  MC_TOKENMIX=gemm timeout 600 python bench.py > gpurun_out/bench_gemm.json 2> gpurun_out/bench_gemm.err; tail -c 600 gpurun_out/bench_gemm.json
fi
