#!/usr/bin/env python
"""One-GPU timing of the BASELINE.json configurations that are not the bench line: config 4's shape (Mixer-B/16,
197 tokens, training step) and config 5 (zero-shot scoring: 1000 class prompts vs 1024 images).  CUDA events, warm-up."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from clip_mixer_b200.clip import CLIP
    from clip_mixer_b200.clip.clip import _MODELS
    from clip_mixer_b200.optim import FusedAdamW
    from clip_mixer_b200.training import FusedTrainStep, synthetic_batch
    from clip_mixer_b200.zeroshot import zeroshot_classifier, zeroshot_logits
    dev = torch.device("cuda", 0)
    out = {}
    # config 5: inference on the B/32 model
    torch.manual_seed(0)
    model = CLIP(**_MODELS["Mixer-B/32"], useTransformer=False, precision="bf16").to(dev).eval()
    images, _ = synthetic_batch(model._cfg, 1024, 7, dev)
    _, prompts = synthetic_batch(model._cfg, 1000, 8, dev)
    with torch.no_grad():
        t_txt = timed(lambda: model.encode_text(prompts), 5)
        W = zeroshot_classifier(model, [prompts[c:c + 1] for c in range(0, 1000, 125)] * 1)   # shape check of the classifier path
        f = model.encode_text(prompts)
        W = (f / f.norm(dim=-1, keepdim=True)).t().contiguous()
        t_img = timed(lambda: zeroshot_logits(model, images, W), 5)
        top5 = zeroshot_logits(model, images, W).topk(5, dim=1).indices
    out["config5_zero_shot"] = {"encode_text_1000_prompts_ms": t_txt, "encode_image_1024_plus_logits_ms": t_img,
                                "images_per_s": 1024 / (t_img * 1e-3), "prompts_per_s": 1000 / (t_txt * 1e-3),
                                "top5_shape": list(top5.shape)}
    del model
    torch.cuda.empty_cache()
    # config 4 shape: B/16 training step on one GPU
    B = int(os.environ.get("B16_BATCH", "256"))
    model = CLIP(**_MODELS["Mixer-B/16"], useTransformer=False, precision="bf16").to(dev).train()
    stepper = FusedTrainStep(model, FusedAdamW(model), None, total_steps=10 ** 6, use_cuda_graph=True)
    images, texts = synthetic_batch(model._cfg, B, 1000, dev)
    for _ in range(3):
        stepper.step(images, texts)
    ms = timed(lambda: stepper.step(images, texts), 10)
    out["config4_shape_B16_train_1gpu"] = {"per_gpu_batch": B, "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3),
                                          "tflops": B / (ms * 1e-3) * 98.169e9 / 1e12, "sm_split": stepper.sm_split,
                                          "loss": float(stepper.loss)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
