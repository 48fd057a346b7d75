#!/bin/bash
# round 2, second 8-GPU call: BASELINE configs[2] (global batch 32768) with the tensor-core head
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
MC_SM_SPLIT=off timeout 300 $TR --master-port 29561 bench.py --gpus 8 --config 3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_8gpu_config3_tc_head.json 2> gpurun_out/r2_bench_8gpu_config3_tc_head.err
tail -c 600 gpurun_out/r2_bench_8gpu_config3_tc_head.json; tail -3 gpurun_out/r2_bench_8gpu_config3_tc_head.err | cut -c1-300
