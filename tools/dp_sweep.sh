#!/bin/bash
# Data-parallel overlap sweep on an N-GPU box: SMs left to NCCL (MC_SM_LIMIT) x NCCL CTA budget.  usage: dp_sweep.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
port=29600
for cfg in "0 0" "140 8" "140 0" "144 4" "132 16"; do
  set -- $cfg
  port=$((port + 1))
  env_args=""
  [ "$1" != "0" ] && env_args="MC_SM_LIMIT=$1"
  [ "$2" != "0" ] && env_args="$env_args NCCL_MAX_CTAS=$2"
  out=$(env $env_args python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$out" | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('MC_SM_LIMIT=$1 NCCL_MAX_CTAS=$2 ->', round(d['value']), 'samples/s', round(d['ms_per_step'], 3), 'ms')" || echo "MC_SM_LIMIT=$1 NCCL_MAX_CTAS=$2 -> failed"
done
