#!/bin/bash
# round 2, final single-GPU verification of the shipped defaults (LayerNorm prologue on): smoke(), the whole GPU suite, the
# default bench line + reference arm, config 3 (micro-batched, one GPU) and config 5 sanity lines, and the ncu launch list
# of one eager step (per-kernel time + DRAM bytes) that profiles/r2_gemm_traffic.json is regenerated from.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
COMMIT=${1:-unknown}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.txt 2>&1; tail -2 gpurun_out/r2h_smoke.txt
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | grep -E "^\[|passed|failed|FAILED|Error|EXEMPT|skipped|losses" | cut -c1-400 > gpurun_out/r2h_gpu_tests.log; tail -3 gpurun_out/r2h_gpu_tests.log
timeout 400 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 600 gpurun_out/r2h_bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/r2h_bench_reference.json
timeout 300 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/r2h_bench_config3_1gpu.json 2> gpurun_out/r2h_bench_config3.err; tail -c 400 gpurun_out/r2h_bench_config3_1gpu.json; tail -3 gpurun_out/r2h_bench_config3.err
timeout 300 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r2h_bench_config5.json 2> gpurun_out/r2h_bench_config5.err; tail -c 400 gpurun_out/r2h_bench_config5.json
python tools/profile_step.py --dump gpurun_out/r2h_gemm_table.txt > gpurun_out/profile_step.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/r2h_ncu_launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r2h_ncu_launches.csv gpurun_out/r2h_gemm_traffic.json $COMMIT > gpurun_out/r2h_launch_summary.txt 2>&1; head -24 gpurun_out/r2h_launch_summary.txt
