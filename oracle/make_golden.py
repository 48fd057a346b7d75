"""TEST INFRASTRUCTURE ONLY -- pins the oracle restatement against the REAL reference module and
writes the golden fixtures under tests/golden/.

Runs only in the build container: it imports /root/reference/training/clip/model.py by file
path (the package __init__ pulls ftfy/azure which are absent).  Nothing at test or
bench time reads /root/reference; the fixtures travel instead.

    python oracle/make_golden.py            # regenerate + verify every fixture

Fixtures
  tiny.pt   full tensors: inputs, embeddings, logits, loss, all gradients, world=1 and world=2
  odd.pt    same, awkward sizes
  S2.pt     BASELINE config-1 shapes, 2+2 layers, batch 4: embeddings, loss, logits, per-gradient
            norm + 64 sampled entries (weights / inputs are re-derived from seeds)
"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mixer_clip_oracle as O  # noqa: E402

REF_MODEL = "/root/reference/training/clip/model.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_clip_model", REF_MODEL)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_step(ref, cfg, sd, image, text, world=1):
    """The reference model + the loss lines of training/training.py:158-168, on one process holding
    the concatenated global batch (equivalent to W ranks with DDP averaging)."""
    model = ref.CLIP(cfg["embed_dim"], cfg["image_resolution"], cfg["vision_layers"], cfg["vision_width"],
                     cfg["vision_patch_size"], cfg["context_length"], cfg["vocab_size"],
                     cfg["transformer_width"], max(1, cfg["transformer_width"] // 64),
                     cfg["transformer_layers"], useTransformer=False)
    model.load_state_dict(sd)
    model.train()
    image_features, text_features, logit_scale = model(image, text)
    N = image_features.shape[0]
    n = N // world
    loss_img = torch.nn.CrossEntropyLoss()
    loss_txt = torch.nn.CrossEntropyLoss()
    total = 0.0
    first = None
    for r in range(world):
        i_loc, t_loc = image_features[r * n:(r + 1) * n], text_features[r * n:(r + 1) * n]
        image_features_gathered = image_features.detach()
        text_features_gathered = text_features.detach()
        logits_per_text = logit_scale * t_loc @ image_features_gathered.t()
        logits_per_image = logit_scale * i_loc @ text_features_gathered.t()
        ground_truth = torch.arange(n, dtype=torch.long) + r * n
        total_loss = (loss_img(logits_per_image, ground_truth) + loss_txt(logits_per_text, ground_truth)) / 2
        total = total + total_loss / world
        if first is None:
            first = (logits_per_image.detach(), logits_per_text.detach())
    total.backward()
    grads = {k: (p.grad.detach() if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
    return dict(image_features=image_features.detach(), text_features=text_features.detach(),
                logit_scale=logit_scale.detach(), loss=total.detach(), logits_per_image=first[0],
                logits_per_text=first[1], grads=grads)


def check_against_oracle(name, refout, cfg, sd, image, text, world):
    got = O.loss_and_grads(sd, image, text, world=world)
    errs = {k: O.l2_rel(got[k], refout[k]) for k in
            ("image_features", "text_features", "logit_scale", "loss", "logits_per_image", "logits_per_text")}
    worst, fails = O.compare_grads(got["grads"], refout["grads"], tol=1e-5)
    print(f"[{name} world={world}] oracle-vs-reference:", {k: f"{v:.2e}" for k, v in errs.items()},
          f"grads worst {worst:.2e} fails {fails}")
    assert max(errs.values()) < 1e-5 and not fails, "oracle restatement disagrees with the reference"
    # closed-form head vs autograd (the formulas the CUDA head kernel implements)
    if world == 1:
        loss_cf, dui, dut, dt = O.head_closed_form(refout["image_features"].double(), refout["text_features"].double(),
                                                    sd["logit_scale"].double())
        assert abs(float(loss_cf) - float(refout["loss"])) < 1e-5
        assert abs(float(dt) - float(refout["grads"]["logit_scale"])) < 1e-4 * max(1.0, abs(float(dt)))


def main():
    ref = load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(os.cpu_count())

    for name, batch in (("tiny", 4), ("odd", 6)):
        cfg = O.CONFIGS[name]
        sd = O.seeded_state_dict(cfg, seed=0)
        image, text = O.synthetic_batch(cfg, batch, seed=1)
        fixture = dict(config=cfg, state_dict=sd, image=image, text=text, world={})
        for world in (1, 2):
            refout = reference_step(ref, cfg, sd, image, text, world)
            check_against_oracle(name, refout, cfg, sd, image, text, world)
            fixture["world"][world] = refout
        torch.save(fixture, os.path.join(out_dir, f"{name}.pt"))
        print("wrote", name, os.path.getsize(os.path.join(out_dir, f"{name}.pt")), "bytes")

    # S2: real widths; weights / inputs come from seeds, only results are stored
    cfg = O.CONFIGS["S2"]
    sd = O.seeded_state_dict(cfg, seed=0)
    image, text = O.synthetic_batch(cfg, 4, seed=1)
    refout = reference_step(ref, cfg, sd, image, text, 1)
    check_against_oracle("S2", refout, cfg, sd, image, text, 1)
    gen = torch.Generator().manual_seed(7)
    gsum = {}
    for k, g in refout["grads"].items():
        flat = g.reshape(-1)
        idx = torch.randint(0, flat.numel(), (min(64, flat.numel()),), generator=gen)
        gsum[k] = dict(norm=float(flat.double().norm()), idx=idx, val=flat[idx].clone())
    fixture = dict(config=cfg, seed_weights=0, seed_batch=1, batch=4,
                   image_features=refout["image_features"], text_features=refout["text_features"],
                   logit_scale=refout["logit_scale"], loss=refout["loss"],
                   logits_per_image=refout["logits_per_image"], logits_per_text=refout["logits_per_text"],
                   grad_summary=gsum)
    torch.save(fixture, os.path.join(out_dir, "S2.pt"))
    print("wrote S2", os.path.getsize(os.path.join(out_dir, "S2.pt")), "bytes")


if __name__ == "__main__":
    main()
