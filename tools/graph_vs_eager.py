#!/usr/bin/env python
"""Bisect of the graph / two-stream discrepancy: tiny config, N optimisation steps on one batch, per-step losses under a
schedule variant.  usage: graph_vs_eager.py {eager1|eager2|graph1|graph2} [--sync] [--precision fp32|bf16] [--steps N]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode")
    ap.add_argument("--sync", action="store_true")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--steps", type=int, default=12)
    a = ap.parse_args()
    from clip_mixer_b200.clip import CLIP
    from clip_mixer_b200.training import FusedTrainStep
    from oracle import mixer_clip_oracle as O
    cfg = O.CONFIGS["tiny"]
    sd = O.seeded_state_dict(cfg, seed=0)
    image, text = O.synthetic_batch(cfg, 8, seed=1)
    image, text = image.cuda(), text.cuda()
    m = CLIP(cfg["embed_dim"], cfg["image_resolution"], cfg["vision_layers"], cfg["vision_width"], cfg["vision_patch_size"],
             cfg["context_length"], cfg["vocab_size"], cfg["transformer_width"], 1, cfg["transformer_layers"],
             useTransformer=False, precision=a.precision)
    m.load_state_dict(sd)
    m = m.cuda().train()
    st = FusedTrainStep(m, total_steps=40, warmup_steps=2, use_cuda_graph=a.mode.startswith("graph"),
                        overlap_towers=a.mode.endswith("2"))
    losses, hy = [], []
    for i in range(a.steps):
        losses.append(st.step(image, text).clone())
        hy.append(st.opt.hyper.clone())
        if a.sync:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    upd = torch.cat([(p.detach().double().cpu() - sd[k].double()).reshape(-1) for k, p in m.named_parameters()])
    print(json.dumps({"mode": a.mode, "sync": a.sync, "precision": a.precision, "sm_split": st.sm_split,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("MC_")},
                      "losses": [round(float(l), 5) for l in losses], "lr": [round(float(h[0]), 8) for h in hy],
                      "upd_norm": float(upd.norm()), "upd_head": [round(float(x), 6) for x in upd[:4]]}), flush=True)


if __name__ == "__main__":
    main()
