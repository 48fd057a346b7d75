#!/bin/bash
# round 2, 2-GPU call: the NCCL multi-rank parity tests that never ran in round 1 + a 2-GPU bench line with the timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_dp2_gpus.txt
timeout 600 python -m pytest tests/test_dp_gpu.py -q -s 2>&1 | grep -E "^\{|passed|failed|skipped|Error" | tail -12 > gpurun_out/r2_dp2_tests.log
for prec in fp32 bf16; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dp_check.py $prec 2>/dev/null | grep "^{" >> gpurun_out/r2_dp_check.jsonl
done
MC_DP_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
cat gpurun_out/r2_dp2_tests.log gpurun_out/r2_dp_check.jsonl; tail -c 1500 gpurun_out/r2_bench_2gpu.json
