#!/bin/bash
# round 2, last 8-GPU call: the bench line of the shipped defaults (LayerNorm prologue on) with the all-reduce timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2h_dp8_gpus.txt
MC_DP_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2h_bench_8gpu.json 2> gpurun_out/r2h_bench_8gpu.err
tail -c 1200 gpurun_out/r2h_bench_8gpu.json
