#!/usr/bin/env python
"""Launched under torchrun (one rank per GPU): W-rank FusedTrainStep on the tiny config vs the oracle on the
concatenated global batch (W ranks with gradient averaging == one process on the global batch).
Rank 0 prints one JSON line with the comparison."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from clip_mixer_b200.clip import CLIP
    from clip_mixer_b200.dp import DataParallel
    from clip_mixer_b200.training import FusedTrainStep
    from oracle import mixer_clip_oracle as O
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    cfg = O.CONFIGS["tiny"]
    sd = O.seeded_state_dict(cfg, seed=0)
    n = 4
    image, text = O.synthetic_batch(cfg, n * world, seed=1)
    model = CLIP(cfg["embed_dim"], cfg["image_resolution"], cfg["vision_layers"], cfg["vision_width"],
                 cfg["vision_patch_size"], cfg["context_length"], cfg["vocab_size"], cfg["transformer_width"], 1,
                 cfg["transformer_layers"], useTransformer=False, precision=precision)
    model.load_state_dict(sd)
    model = model.to(dev).train()
    dp = DataParallel(model, min_bucket_mb=0.01)
    step = FusedTrainStep(model, dp=dp, total_steps=100)
    # run the backward only (no optimizer effect on the gradients): grads are read before the update lands
    sl = slice(rank * n, (rank + 1) * n)
    loss = step.step(image[sl].to(dev), text[sl].to(dev))
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().clone().cpu() for k, p in model.named_parameters()}
    loss_all = loss.clone()
    dist.all_reduce(loss_all)
    # replicas must stay bit-identical: every rank holds the same all-reduced gradients, and the global grad norm (clip
    # coefficient) is reduced in a fixed order (atomics made it rank-dependent)
    for _ in range(3):
        step.step(image[sl].to(dev), text[sl].to(dev))
    torch.cuda.synchronize()
    flat = model._require_store().flat_p
    chk = torch.stack([flat.double().sum(), flat.double().abs().sum(), flat[::9973].double().sum()])
    allchk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allchk, chk)
    replicas_identical = all(torch.equal(allchk[0], c) for c in allchk)
    if rank == 0:
        truth = O.loss_and_grads({k: v.double() for k, v in sd.items()}, image.double(), text, world=world)
        tol = 1e-5 if precision == "fp32" else 2e-2
        worst, fails = O.compare_grads(grads, truth["grads"], tol)
        print(json.dumps({"world": world, "precision": precision, "loss": float(loss_all) / world,
                          "oracle_loss": float(truth["loss"]), "worst_grad_err": worst,
                          "failed": [f[0] for f in fails], "buckets": len(dp.reducer.buckets),
                          "replicas_bit_identical_after_4_steps": replicas_identical}), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
