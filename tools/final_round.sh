#!/bin/bash
# End-of-session verification on one B200: GPU test suite, smoke(), the default bench line, and the ncu launch list of one step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_s3.log 2>&1; tail -3 gpurun_out/gpu_tests_s3.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s3.log 2>&1; tail -1 gpurun_out/smoke_s3.log
timeout 200 python bench.py > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err; tail -c 600 gpurun_out/bench_s3.json
timeout 100 python tools/profile_step.py --dump gpurun_out/gemm_table_s3.txt > gpurun_out/profile_step_s3.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/launches_s3.csv python tools/profile_step.py > gpurun_out/ncu_launches_s3.log 2>&1
tail -n 2 gpurun_out/profile_step_s3.log gpurun_out/ncu_launches_s3.log
