"""GPU parity of the fused training step (FusedTrainStep: forward + fused head + backward + clip + AdamW)
against the oracle: loss, and the weights after two optimisation steps vs torch.optim.AdamW driven by the
oracle's gradients (training/training.py:66-89,144-186).  Also checks that the CUDA-graph replay of the step is
bit-identical to the eager schedule."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(cfg, sd, precision):
    from clip_mixer_b200.clip import CLIP
    m = CLIP(cfg["embed_dim"], cfg["image_resolution"], cfg["vision_layers"], cfg["vision_width"],
             cfg["vision_patch_size"], cfg["context_length"], cfg["vocab_size"], cfg["transformer_width"],
             max(1, cfg["transformer_width"] // 64), cfg["transformer_layers"], useTransformer=False, precision=precision)
    m.load_state_dict(sd)
    return m.to(DEV).train()


def oracle_two_steps(cfg, sd, batches, lrs):
    from oracle import mixer_clip_oracle as O
    from clip_mixer_b200.params import no_decay
    params = {k: torch.nn.Parameter(v.double().clone()) for k, v in sd.items()}
    decay = [p for k, p in params.items() if not no_decay(k, p.ndim)]
    nodecay = [p for k, p in params.items() if no_decay(k, p.ndim)]
    opt = torch.optim.AdamW([{"params": nodecay, "weight_decay": 0.0}, {"params": decay, "weight_decay": 0.2}],
                            lr=5e-4, betas=(0.9, 0.98), eps=1e-6)
    losses = []
    for (image, text), lr in zip(batches, lrs):
        out = O.loss_and_grads({k: p.detach() for k, p in params.items()}, image.double(), text)
        losses.append(float(out["loss"]))
        for k, p in params.items():
            p.grad = out["grads"][k].clone()
        with torch.no_grad():
            params["logit_scale"].data.clamp_(max=100)                     # training.py:178 (non-DDP branch)
        torch.nn.utils.clip_grad_norm_(list(params.values()), 20)          # training.py:181
        for g in opt.param_groups:
            g["lr"] = lr
        opt.step()
    return losses, {k: p.detach() for k, p in params.items()}


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
def test_fused_step_matches_oracle_adamw(precision, tol):
    from clip_mixer_b200.optim import cosine_warmup_lr
    from clip_mixer_b200.training import FusedTrainStep
    from oracle import mixer_clip_oracle as O
    cfg = O.CONFIGS["tiny"]
    sd = O.seeded_state_dict(cfg, seed=0)
    batches = [O.synthetic_batch(cfg, 8, seed=s) for s in (1, 2)]
    lrs = [cosine_warmup_lr(s, 100, 5e-4, 5e-6, 2) for s in (0, 1)]
    ref_losses, ref_params = oracle_two_steps(cfg, sd, batches, lrs)
    model = _model(cfg, sd, precision)
    stepper = FusedTrainStep(model, total_steps=100)
    losses = [float(stepper.step(im.to(DEV), tx.to(DEV))) for im, tx in batches]
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= tol * abs(b), (losses, ref_losses)
    # compare the UPDATE (new - old), the weights themselves would hide errors behind their magnitude
    worst = 0.0
    for k, p in model.named_parameters():
        upd = p.detach().double().cpu() - sd[k].double()
        ref = ref_params[k] - sd[k].double()
        den = float(ref.norm())
        if den < 1e-12:
            continue
        err = float((upd - ref).norm()) / den
        worst = max(worst, err)
    print(f"[train-step/{precision}] losses {losses} vs {ref_losses}; worst update error {worst:.2e}")
    # Adam normalises the gradient: the first updates are ~lr*sign(g), so an element whose gradient error exceeds
    # |g| flips its whole step.  With a fraction f of such elements the update error is ~2*sqrt(f): the fp32 path
    # must stay at 1e-3, the bf16 path (gradient errors ~1e-2) at 0.25.  The precise gradient check is in
    # test_model_parity_gpu.py; this test pins the optimizer wiring (groups, clip, schedule, bias correction).
    assert worst <= (0.25 if precision == "bf16" else 50 * tol), worst


def test_cuda_graph_replay_equals_eager():
    from clip_mixer_b200.training import FusedTrainStep
    from oracle import mixer_clip_oracle as O
    cfg = O.CONFIGS["tiny"]
    sd = O.seeded_state_dict(cfg, seed=0)
    batches = [O.synthetic_batch(cfg, 8, seed=s) for s in (1, 2, 3)]
    outs = []
    for use_graph in (False, True):
        model = _model(cfg, sd, "bf16")
        stepper = FusedTrainStep(model, total_steps=100, use_cuda_graph=use_graph)
        losses = [float(stepper.step(im.to(DEV), tx.to(DEV))) for im, tx in batches]
        outs.append((losses, {k: p.detach().clone() for k, p in model.named_parameters()}))
    (l0, p0), (l1, p1) = outs
    # atomics (split-K, LayerNorm parameter gradients) make the summation order run-dependent: compare closely,
    # not bitwise
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 1e-4 * abs(a), (l0, l1)
    for k in p0:
        d = float((p0[k] - p1[k]).double().norm() / p0[k].double().norm().clamp_min(1e-30))
        assert d <= 1e-3, (k, d)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_micro_batched_step_is_exact(precision, tol):
    """The detached gather makes micro-batching exact: gradients of a step on 8 samples in micro
    batches of 2 equal those of the one-shot step (and the oracle's)."""
    from clip_mixer_b200.training import FusedTrainStep
    from oracle import mixer_clip_oracle as O
    cfg = O.CONFIGS["tiny"]
    sd = O.seeded_state_dict(cfg, seed=0)
    image, text = O.synthetic_batch(cfg, 8, seed=1)
    truth = O.loss_and_grads({k: v.double() for k, v in sd.items()}, image.double(), text)
    model = _model(cfg, sd, precision)
    stepper = FusedTrainStep(model, total_steps=100, micro_batch=2)
    loss = float(stepper.step(image.to(DEV), text.to(DEV)))
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    assert abs(loss - float(truth["loss"])) <= tol * abs(float(truth["loss"]))
    worst, fails = O.compare_grads(grads, truth["grads"], tol)
    cond = float(truth["logits_per_image"].diag().abs().mean())
    fails = [f for f in fails if not (f[0] == "logit_scale" and
                                      abs(float(grads["logit_scale"]) - float(truth["grads"]["logit_scale"])) <= tol * cond)]
    print(f"[micro-batch/{precision}] worst grad err {worst:.2e}")
    assert not fails, fails[:5]


def test_trainer_resumes_from_accelerate_layout_checkpoint(tmp_path, monkeypatch):
    """training.py:106,218-250: a run that saves after 3 of 6 steps and is restarted continues exactly where the
    uninterrupted run went (weights, AdamW moments and step, schedule position, data position)."""
    from clip_mixer_b200.training import Trainer
    from oracle import mixer_clip_oracle as O
    monkeypatch.chdir(tmp_path)
    cfg = O.CONFIGS["tiny"]
    sd = O.seeded_state_dict(cfg, seed=0)
    full = Trainer(_model(cfg, sd, "fp32"), batch_size=8, steps_per_epoch=6)
    assert (full.startEpoch, full.currentStep) == (0, 0)
    ref = full.train()
    assert len(ref) == 6

    first = Trainer(_model(cfg, sd, "fp32"), batch_size=8, steps_per_epoch=6)
    head = []
    for idx, (images, texts) in enumerate(first.trainLoader):
        if idx == 3:
            break
        head.append(float(first.stepper.step(images.to(DEV), texts.to(DEV))))     # the loader yields HOST batches
    first.save_model(0, 3)
    assert sorted(os.listdir("outputs/checkpoints")) == sorted(
        ["epoch.json", "model.safetensors", "optimizer.bin", "random_states_0.pkl", "scheduler.bin"])

    resumed = Trainer(_model(cfg, sd, "fp32"), batch_size=8, steps_per_epoch=6)
    assert (resumed.startEpoch, resumed.currentStep) == (0, 3)
    assert resumed.optimizer.t == 3 and resumed.stepper.sched_step == 3
    tail = resumed.train()
    got = head + tail
    assert len(got) == 6
    for a, b in zip(got, ref):
        assert abs(a - b) <= 1e-4 * abs(b), (got, ref)


@pytest.mark.parametrize("pinned", [False, True])
def test_input_pipeline_delivers_batches_in_order(pinned):
    """training.py:60-62,149,154 (DataLoader(pin_memory) + .to(device)) as the double-buffered InputPipeline: slots are
    reused across many batches, pageable and pinned host sources, prefetch of batch k+1 issued before batch k is used."""
    from clip_mixer_b200.training import InputPipeline
    pipe = InputPipeline(torch.device(DEV))
    g = torch.Generator().manual_seed(0)
    batches = []
    for i in range(7):
        im = torch.randint(0, 256, (5, 3, 32, 32), dtype=torch.uint8, generator=g)
        tx = torch.randint(0, 1000, (5, 12), generator=g)
        if pinned:
            im, tx = im.pin_memory(), tx.pin_memory()
        batches.append((im, tx))
    got = []
    pipe.prefetch(*batches[0])
    for i in range(len(batches)):
        d_im, d_tx = pipe.next()
        if i + 1 < len(batches):
            pipe.prefetch(*batches[i + 1])
        got.append((d_im.clone(), d_tx.clone()))          # "the step": reads the device slot on the compute stream
        torch.cuda._sleep(int(2e6))                       # keep the compute stream busy while the next copy runs
        pipe.release()
    torch.cuda.synchronize()
    for (im, tx), (d_im, d_tx) in zip(batches, got):
        assert torch.equal(d_im.cpu(), im) and torch.equal(d_tx.cpu(), tx)
    assert pipe.h2d_bytes == 5 * 3 * 32 * 32 + 5 * 12 * 8


def test_trainer_runs_its_validators_like_the_reference(tmp_path, monkeypatch):
    """training.py:98-104,205,211-216 + validation.py:142-179: Trainer.validate() drives validator objects; the zero-shot
    validator rebuilds the classifier from the CURRENT weights, scores a loader of (images, target) and reports top-1 /
    top-5 in percent.  Targets are the oracle's own predictions, so both accuracies must be 100 %."""
    from clip_mixer_b200.training import Trainer
    from clip_mixer_b200.zeroshot import ZeroShotValidator
    from oracle import mixer_clip_oracle as O
    monkeypatch.chdir(tmp_path)
    cfg = O.CONFIGS["tiny"]
    sd = O.seeded_state_dict(cfg, seed=0)
    model = _model(cfg, sd, "fp32")
    classes, templates = 6, 3
    prompts = torch.stack([O.synthetic_batch(cfg, templates, seed=200 + c)[1] for c in range(classes)])
    images, _ = O.synthetic_batch(cfg, 10, seed=77)
    cols = []
    for c in range(classes):
        e = O.encode_text(sd, prompts[c])
        e = (e / e.norm(dim=-1, keepdim=True)).mean(0)
        cols.append(e / e.norm())
    f = O.encode_image(sd, images)
    target = ((f / f.norm(dim=-1, keepdim=True)) @ torch.stack(cols, 1)).argmax(1)
    trainer = Trainer(model, batch_size=8, steps_per_epoch=3, epochs=1)
    trainer.validators = [ZeroShotValidator(trainer, prompts, [(images[:6], target[:6]), (images[6:], target[6:])],
                                            classes_per_chunk=4)]
    out = trainer.validate(0)
    assert out[0]["n"] == 10 and out[0]["top1"] == 100.0 and out[0]["top5"] == 100.0
    assert model.training
