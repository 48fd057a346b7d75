#!/usr/bin/env python
"""Where does the bf16 gradient error at the benchmarked batch come from?  Runs the B/32 training step on bench.py's
inputs under one configuration (environment knobs are read by the library at load time, so every variant is its own
process) and prints per-tensor L2-relative gradient errors against the fp32 oracle (cached on disk between variants).
    python tools/parity_probe.py --batch 256 [--eager] [--reference]      # env: MC_TOKENMIX, MC_TM_NO_AUG, MC_LIB ...
TEST INFRASTRUCTURE (uses oracle/)."""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--eager", action="store_true", help="single-stream eager schedule instead of the captured graph")
    ap.add_argument("--no-graph", action="store_true", help="two-stream schedule, not captured")
    ap.add_argument("--single-stream", action="store_true", help="captured graph of the single-stream schedule")
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--reference", action="store_true", help="the reference's own bf16-autocast path instead of ours")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import test_bench_path_gpu as T
    from oracle import mixer_clip_oracle as O
    cache = f"/tmp/parity_truth_{a.batch}.pt"
    if os.path.exists(cache):
        images_u8, texts, img, truth = torch.load(cache, weights_only=False)
    else:
        images_u8, texts, img, truth = T._bench_batch_and_truth(a.batch)
        torch.save((images_u8, texts, img, truth), cache)
    cfg, sd, model = T._b32("bf16")
    if a.reference:
        loss, grads = T._reference_bf16_autocast_grads(img, texts, sd)
    else:
        from clip_mixer_b200.training import FusedTrainStep
        st = FusedTrainStep(model, total_steps=10 ** 6, use_cuda_graph=not (a.eager or a.no_graph),
                            overlap_towers=not (a.eager or a.single_stream))
        sd0 = {k: p.detach().clone() for k, p in model.named_parameters()}
        for rep in range(a.repeat):
            if rep:                                  # same weights again: run-to-run spread of one configuration
                with torch.no_grad():
                    for k, p in model.named_parameters():
                        p.copy_(sd0[k])
                model.mark_weights_dirty()
                if model._store.flat_w16 is not None:
                    model._store.refresh_mirror(force=True)
            loss = float(st.step(images_u8.to("cuda:0"), texts.to("cuda:0")))
            torch.cuda.synchronize()
            grads = {k: p.grad.detach().float().cpu().clone() for k, p in model.named_parameters()}
            if rep + 1 < a.repeat:
                e_, m_, w_ = T._grad_errors(grads, truth)
                print(json.dumps({"tag": a.tag, "rep": rep, "sm_split": st.sm_split, "median": m_, "whole_model": w_}), flush=True)
    errs, median, whole = T._grad_errors(grads, truth)
    by = {}
    for e, k in errs:
        key = ("img." if k.startswith("visual.") else "txt.") + k.split(".")[-2] + "." + k.split(".")[-1] if "mixBlocks" in k else k
        by.setdefault(key, []).append(e)
    groups = {k: round(sum(v) / len(v), 4) for k, v in sorted(by.items())}
    print(json.dumps({"tag": a.tag, "batch": a.batch, "reference": a.reference, "eager": a.eager,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("MC_")},
                      "sm_split": None if a.reference else st.sm_split, "loss_rel": abs(loss - float(truth["loss"])) / float(truth["loss"]), "worst": errs[0][0], "median": median,
                      "whole_model": whole, "mean_by_kind": groups}), flush=True)


if __name__ == "__main__":
    main()
