"""Checkpoint directories: the optimizer file must load into a real torch.optim.AdamW built the way
the reference builds it (training/training.py:66-82) and round-trip bit-exactly through the flat moment buffers."""
import json
import os
import types

import torch

from clip_mixer_b200 import checkpoint
from clip_mixer_b200.clip import CLIP
from clip_mixer_b200.params import ParamStore, no_decay
from oracle import mixer_clip_oracle as O


def _tiny():
    cfg = O.CONFIGS["odd"]      # 50-ish token shapes whose row pitches are padded in the flat layout
    torch.manual_seed(0)
    m = CLIP(cfg["embed_dim"], cfg["image_resolution"], cfg["vision_layers"], cfg["vision_width"],
             cfg["vision_patch_size"], cfg["context_length"], cfg["vocab_size"], cfg["transformer_width"], 1,
             cfg["transformer_layers"], useTransformer=False)
    buckets, _ = m._flat_order()
    shapes = {n: tuple(p.shape) for n, p in m.named_parameters()}
    store = ParamStore(shapes, [n for b in buckets for n in b], "cpu", buckets)
    return m, store


def _reference_adamw(model):
    """training.py:66-82, verbatim grouping."""
    exclude = lambda n, p: p.ndim < 2 or "bn" in n or "ln" in n or "bias" in n or "logit_scale" in n
    named = list(model.named_parameters())
    gain_or_bias = [p for n, p in named if exclude(n, p) and p.requires_grad]
    rest = [p for n, p in named if not exclude(n, p) and p.requires_grad]
    return torch.optim.AdamW([{"params": gain_or_bias, "weight_decay": 0.0}, {"params": rest, "weight_decay": 0.2}],
                             lr=5e-4, betas=(0.9, 0.98), eps=1e-6)


def _fake_optimizer(store, t=7):
    g = torch.Generator().manual_seed(3)
    opt = types.SimpleNamespace(store=store, m=torch.zeros(store.total), v=torch.zeros(store.total), t=t, lr=5e-4,
                                betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2)
    for s in store.slots.values():      # moments live in the valid elements only; pads stay zero
        s.view(opt.m).copy_(torch.randn(s.shape, generator=g))
        s.view(opt.v).copy_(torch.rand(s.shape, generator=g))
    return opt


def test_param_groups_follow_the_reference_filter():
    model, _ = _tiny()
    named = [(n, tuple(p.shape)) for n, p in model.named_parameters()]
    g0, g1 = checkpoint.reference_param_groups(named)
    ref = _reference_adamw(model)
    assert [len(g["params"]) for g in ref.param_groups] == [len(g0), len(g1)]
    assert "logit_scale" in g0 and "visual.class_embedding" in g0 and "token_embedding.weight" in g1
    assert all(no_decay(n, len(dict(named)[n])) for n in g0)


def test_optimizer_state_loads_into_reference_adamw_and_round_trips():
    model, store = _tiny()
    opt = _fake_optimizer(store)
    named = [(n, tuple(p.shape)) for n, p in model.named_parameters()]
    sd = checkpoint.flat_to_torch_adamw(store, named, opt.m, opt.v, opt.t, opt.lr)
    ref = _reference_adamw(model)
    ref.load_state_dict(sd)                                    # the reference's optimizer accepts the file
    params = dict(model.named_parameters())
    g0, g1 = checkpoint.reference_param_groups(named)
    for name in (g0[0], g0[-1], g1[0], g1[-1], "visual.transformer.mixBlocks.0.token_mix_seq.lin1.weight"):
        st = ref.state[params[name]]
        assert torch.equal(st["exp_avg"], store.slots[name].view(opt.m))
        assert torch.equal(st["exp_avg_sq"], store.slots[name].view(opt.v))
        assert float(st["step"]) == 7.0
    # and what the reference's optimizer saves comes back into the flat buffers bit-exactly
    m2, v2 = torch.full((store.total,), 9.0), torch.full((store.total,), 9.0)
    t2 = checkpoint.torch_adamw_to_flat(ref.state_dict(), store, named, m2, v2)
    assert t2 == 7 and torch.equal(m2, opt.m) and torch.equal(v2, opt.v)


def test_directory_round_trip(tmp_path):
    model, store = _tiny()
    opt = _fake_optimizer(store, t=11)
    path = str(tmp_path / "checkpoints")
    checkpoint.save_state(path, model, opt, sched_step=11, total_steps=100, rank=0)
    checkpoint.write_epoch_json(path, 2, 5)
    files = set(os.listdir(path))
    assert {"optimizer.bin", "scheduler.bin", "random_states_0.pkl", "epoch.json"} <= files
    assert "model.safetensors" in files or "pytorch_model.bin" in files
    assert json.load(open(os.path.join(path, "epoch.json"))) == {"epoch": 2, "step": 5}

    model2, store2 = _tiny()
    with torch.no_grad():
        for p in model2.parameters():
            p.add_(1.0)
    opt2 = types.SimpleNamespace(store=store2, m=torch.ones(store2.total), v=torch.ones(store2.total), t=0, lr=5e-4,
                                 betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2)
    step = checkpoint.load_state(path, model2, opt2)
    assert step == 11 and opt2.t == 11
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k
    assert torch.equal(opt2.m, opt.m) and torch.equal(opt2.v, opt.v)
    assert checkpoint.read_epoch_json(path) == {"epoch": 2, "step": 5}
    # scheduler.bin: full key set of CosineAnnealingWarmupRestarts.state_dict(); values follow cosine_warmup_lr
    from clip_mixer_b200.optim import cosine_warmup_lr
    sched = torch.load(os.path.join(path, "scheduler.bin"), weights_only=False)
    assert {"first_cycle_steps", "cycle_mult", "base_max_lr", "max_lr", "min_lr", "warmup_steps", "gamma", "cur_cycle_steps",
            "cycle", "step_in_cycle", "base_lrs", "last_epoch", "_step_count", "_last_lr"} <= set(sched)
    assert sched["last_epoch"] == 11 and sched["step_in_cycle"] == 11 and sched["cycle"] == 0
    assert abs(sched["_last_lr"][0] - cosine_warmup_lr(11, 100, 5e-4, 5e-6, 2)) < 1e-12
    # random_states_<rank>.pkl: written with torch.save, accelerate's keys, and the torch RNG state round-trips
    rng = torch.load(os.path.join(path, "random_states_0.pkl"), weights_only=False)
    assert {"step", "random_state", "numpy_random_seed", "torch_manual_seed"} <= set(rng)
    torch.set_rng_state(rng["torch_manual_seed"])
    a = torch.rand(3)
    torch.set_rng_state(rng["torch_manual_seed"])
    assert torch.equal(a, torch.rand(3))


def test_ddp_prefixed_weights_and_missing_directory(tmp_path):
    model, store = _tiny()
    opt = _fake_optimizer(store)
    path = str(tmp_path / "ck")
    checkpoint.save_state(path, model, opt, sched_step=3, safe_serialization=False)
    sd = torch.load(os.path.join(path, "pytorch_model.bin"))
    torch.save({"module." + k: v for k, v in sd.items()}, os.path.join(path, "pytorch_model.bin"))
    model2, store2 = _tiny()
    opt2 = _fake_optimizer(store2, t=0)
    assert checkpoint.load_state(path, model2, opt2) == 3
    try:
        checkpoint.load_state(str(tmp_path / "nothing"), model2, opt2)
        raise AssertionError("a missing checkpoint must raise (the trainer then starts from scratch)")
    except FileNotFoundError:
        pass
