rm -f gpurun_out/run12.log
port=29630
for cfg in "MC_SM_SPLIT=78,62 NCCL_MAX_CTAS=8" "MC_SM_SPLIT=80,64 NCCL_MAX_CTAS=4" "MC_SM_SPLIT=82,66 NCCL_MAX_CTAS=8"; do
  port=$((port + 1))
  out=$(env $cfg timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 15 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$out" | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('$cfg ->', round(d['value']), 'samples/s', round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['ms_per_step'], 3))" >> gpurun_out/run12.log 2>&1 || echo "$cfg failed" >> gpurun_out/run12.log
done
cat gpurun_out/run12.log
