#!/bin/bash
# round 2, third 8-GPU call: all-reduce bucket-size sweep (the first sweep: 8 MB 16.31 ms, 32 MB 16.13, 128 MB 15.80)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
i=0
for mb in 64 96 128 256; do
  i=$((i+1))
  MC_DP_BUCKET_MB=$mb MC_SM_SPLIT=84,64 MC_DP_TRACE=1 timeout 150 $TR --master-port 2958$i bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_8gpu_bucket$mb.json 2>/dev/null
  python - gpurun_out/r2_bench_8gpu_bucket$mb.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); tl=d.get('dp_timeline_rank0') or {}
    print(sys.argv[1], round(d['value']), 'samples/s', round(d['ms_per_step'],3), 'ms; buckets', len(tl.get('buckets',[])), 'tail', tl.get('exposed_tail_ms'), 'busy', tl.get('allreduce_busy_ms'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
