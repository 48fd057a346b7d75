#!/usr/bin/env python
"""Benchmark of the Mixer-CLIP training hot path (BASELINE.json metric: train samples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                          # the reference algorithm on the host CPUs

Workload (config.workload): BASELINE.json configs[1] -- Mixer-CLIP B/32-size (111M parameters), bf16
tensor-core training step (forward + contrastive loss + backward + grad clip + AdamW), 256 samples per GPU,
synthetic 224x224 uint8 images and 77-token texts.  With N > 1 (torchrun, one rank per GPU) every rank keeps
256 samples (weak scaling), features are all-gathered for the global-batch loss and gradients are
bucket-all-reduced behind the backward pass.

One JSON line on stdout (rank 0).  `value`: whole-job samples/s with inputs resident in HBM; `e2e`: the same
step fed from pinned HOST buffers with the H2D copies and a D2H read of the loss inside the timed region;
`roofline`: the tcgen05 GEMM engine (dominant kernel) timed per launch with CUDA events in a separate
instrumented pass; `cpu_baseline`: the oracle port timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

PER_GPU_BATCH = 256
MODEL = "B32"
TRAIN_GFLOP_PER_SAMPLE = 32.166          # GEMM FLOPs of one training sample (3x forward)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--per-gpu-batch", type=int, default=PER_GPU_BATCH)
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step into a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md): sampled DURING the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_step_time(model_name, batch, iters, warmup=1):
    from oracle import mixer_clip_oracle as O
    torch.set_num_threads(os.cpu_count())
    cfg = O.CONFIGS[model_name]
    sd = O.seeded_state_dict(cfg, seed=0)
    image, text = O.synthetic_batch(cfg, batch, seed=1)
    times = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        O.loss_and_grads(sd, image, text)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 8
    times = cpu_step_time(args.model, batch, max(1, args.steps), max(1, min(args.warmup, 2)))
    t = sum(times) / len(times)
    value = batch / t
    cores = os.cpu_count()
    sample = f"{args.model} 12+12 layers, batch {batch}, forward + loss + backward, fp32, {len(times)} timed iterations"
    line = {"impl": "reference", "metric": "mixer_clip_train_samples_per_sec", "value": value, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": max(1, min(args.warmup, 2)), "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"Mixer-CLIP {args.model}-size (BASELINE.json configs[1]) training step: forward + contrastive "
                        f"loss + backward + clip_grad_norm + AdamW, {args.per_gpu_batch} samples per GPU",
            "global_batch": args.per_gpu_batch * world, "image": "224x224 uint8", "text_tokens": 77,
            "parallelism": f"dp{world}", "l2_hygiene": "per-step working set (~10 GB of activations) >> 126 MB L2",
            "cuda_graph": not args.no_graph, "streams": "image and text towers on two streams"}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback of the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from clip_mixer_b200 import ops
    from clip_mixer_b200.clip import CLIP
    from clip_mixer_b200.clip.clip import _MODELS
    from clip_mixer_b200.dp import DataParallel
    from clip_mixer_b200.optim import FusedAdamW
    from clip_mixer_b200.training import FusedTrainStep, synthetic_batch

    name = {"B32": "Mixer-B/32", "B16": "Mixer-B/16", "S": "Mixer-S/32"}[args.model]
    torch.manual_seed(0)
    model = CLIP(**_MODELS[name], useTransformer=False, precision=args.precision).to(dev).train()
    dp = DataParallel(model) if world > 1 else None
    opt = FusedAdamW(model)
    stepper = FusedTrainStep(model, opt, dp, total_steps=10 ** 6, use_cuda_graph=not args.no_graph)
    B = args.per_gpu_batch
    images, texts = synthetic_batch(model._cfg, B, 1000 + rank, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # launches of OUR kernels per step (counted once on an eager step; graph replays repeat them)
    ops.reset_launch_count()
    for _ in range(args.warmup):
        stepper.step(images, texts)
    barrier()
    # single-stream, eager twin of the step: counts launches and carries the per-kernel CUDA-event timing
    eager = FusedTrainStep(model, opt, dp, total_steps=10 ** 6, use_cuda_graph=False, overlap_towers=False)
    ops.reset_launch_count()
    eager.step(images, texts)
    launches_per_step = ops.launch_count()
    barrier()

    # ---- value: inputs resident in HBM ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        stepper.step(images, texts)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    loss_value = float(stepper.loss)

    # ---- e2e: pinned host buffers -> H2D, step, D2H of the loss, every step ----
    h_images = images.cpu().pin_memory()
    h_texts = texts.cpu().pin_memory()
    h_loss = torch.zeros(1).pin_memory()
    d_images, d_texts = torch.empty_like(images), torch.empty_like(texts)
    for _ in range(2):
        d_images.copy_(h_images, non_blocking=True)
        d_texts.copy_(h_texts, non_blocking=True)
        stepper.step(d_images, d_texts)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        d_images.copy_(h_images, non_blocking=True)
        d_texts.copy_(h_texts, non_blocking=True)
        loss = stepper.step(d_images, d_texts)
        h_loss.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the loss value is consumed on the host every step
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- roofline of the dominant kernel: per-launch CUDA-event timing of the tcgen05 GEMM engine ----
    roof = None
    if args.precision == "bf16":
        ops.enable_gemm_timing(True)
        # park the GPU on a spin kernel while the host enqueues the whole instrumented step, so that the event
        # pairs bracket kernel execution only (an eager step is host-bound: ~460 launches + 2 events each)
        torch.cuda._sleep(int(1.2e8))
        eager.step(images, texts)
        torch.cuda.synchronize()
        recs = ops.collect_gemm_timing()
        ops.enable_gemm_timing(False)
        tc = [r for r in recs if r["engine"] == "tc"]
        t_tc = sum(r["ms"] for r in tc) * 1e-3
        fl_tc = sum(r["flops"] for r in tc)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PF sustained (of fallback)"
        ach = fl_tc / t_tc / 1e12 if t_tc > 0 else 0.0
        big = [r for r in tc if r["flops"] > 2e10]
        t_big = sum(r["ms"] for r in big) * 1e-3
        traffic, traffic_src = None, None
        try:   # DRAM bytes of the same launch set (one step), from the committed ncu launch list of this workload
            from clip_mixer_b200.engine import fused_token_mix_enabled
            names = ["r1s3_gemm_traffic.json", "r1s2_gemm_traffic.json"] if fused_token_mix_enabled() else ["r1_final_gemm_traffic.json"]
            tj = json.load(open(next(f for f in (os.path.join(ROOT, "profiles", n) for n in names) if os.path.exists(f))))
            if args.model == "B32" and B == 256:
                traffic, traffic_src = tj["dram_bytes_per_step"], tj["source"]
        except Exception:
            pass
        # second kernel family: the fused token-mixing kernels (HBM bound at 50 / 77 tokens).  Algorithmic bytes per
        # launch: fwd 10*P*D*B (u bf16 in, x fp32 in, y fp32 out), dgrad 8*P*D*B, wgrad 4*P*D*B
        tm_roof = None
        tm = [r for r in recs if r["engine"].startswith("token_mix")]
        if tm:
            per = {"token_mix_fwd": (10, 2), "token_mix_dgrad": (8, 3), "token_mix_wgrad": (4, 4)}
            tm_bytes = sum(per[r["engine"]][0] * r["K"] * r["N"] * (r["batch"] // per[r["engine"]][1]) for r in tm)
            tm_t = sum(r["ms"] for r in tm) * 1e-3
            hbm = peaks.get("hbm_gbs", 6650.0)
            kinds = {}
            for k in per:
                rr = [r for r in tm if r["engine"] == k]
                if rr:
                    kb = sum(per[k][0] * r["K"] * r["N"] * (r["batch"] // per[k][1]) for r in rr)
                    kt = sum(r["ms"] for r in rr) * 1e-3
                    kinds[k] = {"launches": len(rr), "ms": kt * 1e3, "GBps": kb / kt / 1e9}
            tm_roof = {"bound": "hbm", "kernel": "token_mix_kernel<fwd|dgrad|wgrad> (fused token-mixing MLP, all launches of one step)",
                       "achieved": tm_bytes / tm_t / 1e9, "peak": hbm, "unit": "GB/s", "frac": tm_bytes / tm_t / 1e9 / hbm,
                       "launches": len(tm), "ms_per_step": tm_t * 1e3, "algorithmic_MB_per_step": tm_bytes / 1e6,
                       "tensor_TFLOPs": sum(r["flops"] for r in tm) / tm_t / 1e12, "by_kind": kinds,
                       "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)"}
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 GEMM engine, all launches of one step)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "traffic_unit": "DRAM bytes per step over the same launches", "traffic_source": traffic_src,
                "peak_source": peak_src, "launches": len(tc), "gemm_ms_per_step": t_tc * 1e3,
                "algorithmic_gflop_per_step": fl_tc / 1e9,
                "channel_mix_only": {"achieved": (sum(r["flops"] for r in big) / t_big / 1e12) if t_big else None,
                                     "launches": len(big), "ms": t_big * 1e3},
                "token_mix": tm_roof}

    # ---- reductions over ranks ----
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        total = B * world * args.steps
        value = total / (ms * 1e-3)
        e2e = total / (ms_e2e * 1e-3)
        line = {"metric": "mixer_clip_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": workload_config(args, world), "clocks": clocks,
                "e2e": {"value": e2e, "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": int(h_images.numel() * h_images.element_size() + h_texts.numel() * h_texts.element_size()),
                        "d2h_bytes_per_step": 4},
                "gpu_launches": launches_per_step * args.steps,
                "launches_per_step": launches_per_step,
                "sm_split": {"image_text_sms": stepper.sm_split,
                             "trials_ms": [[list(c) if c else None, round(t, 3)] for c, t in stepper.sm_split_trials]},
                "model_flops_utilisation": value / world * TRAIN_GFLOP_PER_SAMPLE * 1e9 / 1e12 /
                (json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1400.0)
                 if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0) if args.model == "B32" else None,
                "loss": loss_value, "roofline": roof}
        if not args.no_cpu_baseline and world == 1:
            times = cpu_step_time(args.model, 8, 3, 1)
            t = sum(times) / len(times)
            line["cpu_baseline"] = {"value": 8 / t, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{args.model} batch 8, forward + loss + backward on the host (oracle port, "
                                              f"torch fp32, {os.cpu_count()} threads), 3 timed iterations, {t * 1e3:.0f} ms each"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL communicators referenced by a captured CUDA graph can stall the interpreter's teardown:
        # synchronise, meet at a barrier, flush, and leave without running destructors.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
