#!/bin/bash
# round 2, 8-GPU A/B: bucket size for the exposed end of each tower's backward (MC_DP_TAIL_MB)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tail in 64 16 32; do
MC_DP_TAIL_MB=$tail MC_DP_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2955$((tail % 10)) bench.py --gpus 8 --steps 20 --warmup 3 --no-eager-baseline > gpurun_out/r2h_bench_8gpu_tail${tail}MB.json 2> gpurun_out/r2h_bench_8gpu_tail.err
python - $tail <<'PY'
import json, sys
d = None
for line in open(f"gpurun_out/r2h_bench_8gpu_tail{sys.argv[1]}MB.json"):
    if line.startswith("{"):
        d = json.loads(line)
t = d["dp_timeline_rank0"]
print("tail", sys.argv[1], "MB:", round(d["ms_per_step"], 3), "ms", round(d["value"]), "samples/s; split", d["sm_split"]["image_text_sms"],
      "exposed", t["exposed_tail_ms"], "busy", t["allreduce_busy_ms"], "buckets", [b["MB"] for b in t["buckets"]])
PY
done
