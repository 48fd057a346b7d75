#!/usr/bin/env python
"""Runs the bench workload eagerly (no CUDA graph) and brackets ONE training step with
cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees exactly one step.  Also dumps the
per-launch CUDA-event timing of every GEMM of that step (shape, operand majors, ms, TFLOP/s)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="Mixer-B/32")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--dump", default="")
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    from clip_mixer_b200 import ops
    from clip_mixer_b200.clip import CLIP
    from clip_mixer_b200.clip.clip import _MODELS
    from clip_mixer_b200.training import FusedTrainStep, synthetic_batch
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = CLIP(**_MODELS[args.model], useTransformer=False).to(dev).train()
    stepper = FusedTrainStep(model, use_cuda_graph=False, overlap_towers=False)
    images, texts = synthetic_batch(model._cfg, args.batch, 1000, dev)
    for _ in range(args.warmup):
        stepper.step(images, texts)
    torch.cuda.synchronize()
    if args.dump:
        ops.enable_gemm_timing(True)
        torch.cuda._sleep(int(1.2e8))      # let the host run ahead: event pairs then bracket kernel time only
        stepper.step(images, texts)
        recs = ops.collect_gemm_timing()
        ops.enable_gemm_timing(False)
        agg = {}
        for r in recs:
            key = (r["engine"], r["M"], r["N"], r["K"], r["batch"], r["tag"])
            a = agg.setdefault(key, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += r["ms"]
            a[2] += r["flops"]
        rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
        with open(args.dump, "w") as f:
            f.write(f"{'engine':6} {'M':>7} {'N':>6} {'K':>6} {'batch':>5} {'tag':12} {'n':>4} {'ms_total':>9} {'us_each':>9} {'TFLOP/s':>8}\n")
            for (eng, M, N, K, b, tag), (n, ms, fl) in rows:
                f.write(f"{eng:6} {M:7d} {N:6d} {K:6d} {b:5d} {tag:12} {n:4d} {ms:9.3f} {ms / n * 1e3:9.1f} {fl / ms / 1e9:8.1f}\n")
            f.write(f"total gemm ms {sum(v[1] for v in agg.values()):.3f}\n")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    stepper.step(images, texts)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(json.dumps({"eager_step_ms": e0.elapsed_time(e1), "loss": float(stepper.loss)}))


if __name__ == "__main__":
    main()
