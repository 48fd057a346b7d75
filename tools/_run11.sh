NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 15 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_8gpu_s3.err | tail -1 > gpurun_out/bench_8gpu_s3.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_8gpu_s3.json").read())
print("8 GPUs:", round(d["value"]), "samples/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["ms_per_step"], 3), d.get("sm_split"), d["clocks"])
PY
tail -5 gpurun_out/bench_8gpu_s3.err
