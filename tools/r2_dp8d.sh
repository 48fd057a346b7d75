#!/bin/bash
# round 2, final 8-GPU run with pure defaults (what the driver's scaling bench will launch), timeline recorded
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
MC_DP_TRACE=1 timeout 240 $TR --master-port 29591 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_8gpu_final.json 2> gpurun_out/r2_bench_8gpu_final.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu_final.json 2>/dev/null
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29593 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_bench_4gpu_final.json 2>/dev/null
for f in gpurun_out/r2_bench_2gpu_final.json gpurun_out/r2_bench_4gpu_final.json gpurun_out/r2_bench_8gpu_final.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); tl=d.get('dp_timeline_rank0') or {}
    print(sys.argv[1], round(d['value']), 'samples/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['ms_per_step'],3), 'split', d['sm_split']['image_text_sms'], 'buckets', len(tl.get('buckets',[])), 'tail', tl.get('exposed_tail_ms'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -2 gpurun_out/r2_bench_8gpu_final.err | cut -c1-200
