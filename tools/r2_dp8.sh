#!/bin/bash
# round 2, 8-GPU call: rank-8 NCCL parity (tools/dp_check.py), weak-scaling bench with the measured all-reduce timeline,
# BASELINE configs[2] (global batch 32768 = 8 x 4096, one shot), opt-in bf16 gradient wire
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for prec in fp32 bf16; do
  timeout 120 $TR --master-port 29551 tools/dp_check.py $prec 2>/dev/null | grep "^{" >> gpurun_out/r2_dp_check8.jsonl
done
MC_DP_TRACE=1 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 240 $TR --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
grep -m5 -i "nvls\|algo\|Using network\|channels" gpurun_out/r2_bench_8gpu.err | cut -c1-200 > gpurun_out/r2_nccl_info.txt
MC_DP_BF16=1 timeout 200 $TR --master-port 29553 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_8gpu_bf16wire.json 2> /dev/null
MC_SM_SPLIT=off timeout 300 $TR --master-port 29554 bench.py --gpus 8 --config 3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_8gpu_config3.json 2> gpurun_out/r2_bench_8gpu_config3.err
cat gpurun_out/r2_dp_check8.jsonl; for f in gpurun_out/r2_bench_8gpu.json gpurun_out/r2_bench_8gpu_bf16wire.json gpurun_out/r2_bench_8gpu_config3.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value']), 'samples/s', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['ms_per_step'],2), 'tail', (d.get('dp_timeline_rank0') or {}).get('exposed_tail_ms'))
except Exception as e:
    print(sys.argv[1], 'no line', e)
PY
done
tail -3 gpurun_out/r2_bench_8gpu_config3.err
