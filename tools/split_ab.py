"""Interleaved A/B of SM-split settings on the 1-GPU training step (one process, one model, one graph per setting;
blocks of replays alternate so that clock / power drift hits every setting alike).
usage: python tools/split_ab.py [off 84,64 82,66 ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from clip_mixer_b200.clip import CLIP
    from clip_mixer_b200.clip.clip import _MODELS
    from clip_mixer_b200.optim import FusedAdamW
    from clip_mixer_b200.training import FusedTrainStep, synthetic_batch
    settings = sys.argv[1:] or ["off", "84,64", "82,66"]
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = CLIP(**_MODELS["Mixer-B/32"], useTransformer=False, precision="bf16").to(dev).train()
    opt = FusedAdamW(model)
    images, texts = synthetic_batch(model._cfg, 256, 1000, dev)
    steppers = []
    for s in settings:
        os.environ["MC_SM_SPLIT"] = s
        st = FusedTrainStep(model, opt, None, total_steps=10 ** 6, use_cuda_graph=True)
        for _ in range(3):
            st.step(images, texts)
        steppers.append(st)
    torch.cuda.synchronize()
    blocks, reps = int(os.environ.get("BLOCKS", "5")), int(os.environ.get("REPS", "25"))
    res = {s: [] for s in settings}
    for b in range(blocks):
        for s, st in zip(settings, steppers):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                st.step(images, texts)
            e1.record()
            e1.synchronize()
            res[s].append(e0.elapsed_time(e1) / reps)
    for s in settings:
        v = res[s]
        print(f"MC_SM_SPLIT={s:8s} ms/step per block: {' '.join(f'{x:.3f}' for x in v)}   mean {sum(v) / len(v):.3f}  min {min(v):.3f}")


if __name__ == "__main__":
    main()
