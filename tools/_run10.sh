timeout 600 python -m pytest tests/test_dp_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/run10.log
for cfg in "MC_SM_SPLIT=auto" "MC_SM_SPLIT=off"; do
env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_2gpu_s3.err | tail -1 > gpurun_out/bench_2gpu_s3_$cfg.json
python - "$cfg" <<'PY' >> gpurun_out/run10.log
import json, sys
d = json.loads(open(f"gpurun_out/bench_2gpu_s3_{sys.argv[1]}.json").read())
print(sys.argv[1], round(d["value"]), "samples/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["ms_per_step"], 3), d.get("sm_split"))
PY
done
cat gpurun_out/run10.log
