#!/bin/bash
# round 2, GPU call 10: tensor-core head (tests + timing at n=4096, N=32768), configs 4 and 5, sanitizer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bench_path_gpu.py -q -s -k "head_at_global" 2>&1 | grep -E "^\[|passed|failed|FAILED|Error" | cut -c1-300 > gpurun_out/r2_head_tests.log
cat gpurun_out/r2_head_tests.log
timeout 300 python - > gpurun_out/r2_head_timing.txt 2>&1 <<'PY'
import os, sys, math, torch
sys.path.insert(0, os.getcwd())
from clip_mixer_b200 import ops
dev = "cuda:0"
E = 512
for (n, N) in ((4096, 32768), (2048, 32768), (256, 2048), (256, 256)):
    g = torch.Generator().manual_seed(1)
    ua = torch.nn.functional.normalize(torch.randn(N, E, generator=g), dim=1).to(dev)
    ta = torch.nn.functional.normalize(torch.randn(N, E, generator=g), dim=1).to(dev)
    t = torch.tensor([math.log(1 / 0.07)], device=dev)
    for tc in ("0", "1"):
        os.environ["MC_HEAD_TC"] = tc
        ws = torch.empty(ops.head_workspace_bytes(n, N, E) // 4, device=dev)
        loss, dls = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
        dui, dut = torch.empty(n, E, device=dev), torch.empty(n, E, device=dev)
        def run():
            ops.head_fwd_bwd(ua[:n], ta[:n], ua, ta, t, n, N, E, 0, 1.0, loss, dui, dut, dls, ws)
        try:
            for _ in range(2): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            fl = 8.0 * n * N * E
            byts = 4.0 * (2 * n * E + 2 * N * E + 2 * n * E)
            print(f"head n={n} N={N} MC_HEAD_TC={tc}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s algorithmic  minimal-bytes {byts / 1e6:.0f} MB -> {byts / ms / 1e6:.1f} GB/s", flush=True)
        except Exception as ex:
            print(f"head n={n} N={N} MC_HEAD_TC={tc}: {type(ex).__name__} {str(ex)[:200]}", flush=True)
PY
cat gpurun_out/r2_head_timing.txt
timeout 500 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config4.json 2> gpurun_out/r2_bench_config4.err
timeout 300 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2_bench_config5.json 2> gpurun_out/r2_bench_config5.err
timeout 300 python bench.py --config 5 --templates 80 --steps 3 --warmup 1 > gpurun_out/r2_bench_config5_80.json 2> gpurun_out/r2_bench_config5_80.err
for f in gpurun_out/r2_bench_config4.json gpurun_out/r2_bench_config5.json gpurun_out/r2_bench_config5_80.json; do tail -c 900 $f; echo; done
tail -3 gpurun_out/r2_bench_config4.err gpurun_out/r2_bench_config5.err | cut -c1-300
