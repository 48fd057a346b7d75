#!/bin/bash
# round 2, second 8-GPU call: BASELINE configs[2] (global batch 32768) with the tensor-core head
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
MC_SM_SPLIT=off timeout 300 $TR --master-port 29561 bench.py --gpus 8 --config 3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_8gpu_config3_tc_head.json 2> gpurun_out/r2_bench_8gpu_config3_tc_head.err
for mb in 8 32 128; do
  MC_DP_BUCKET_MB=$mb MC_SM_SPLIT=84,64 MC_DP_TRACE=1 timeout 200 $TR --master-port 2957$((mb % 10)) bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_8gpu_bucket$mb.json 2>/dev/null
  python - gpurun_out/r2_bench_8gpu_bucket$mb.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); tl=d.get('dp_timeline_rank0') or {}
    print(sys.argv[1], round(d['value']), 'samples/s', round(d['ms_per_step'],3), 'ms; buckets', len(tl.get('buckets',[])), 'tail', tl.get('exposed_tail_ms'), 'busy', tl.get('allreduce_busy_ms'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -c 600 gpurun_out/r2_bench_8gpu_config3_tc_head.json; tail -3 gpurun_out/r2_bench_8gpu_config3_tc_head.err | cut -c1-300
