"""Checkpoint directories in the layout the reference trains with.

The reference saves and resumes through ``accelerator.save_state(dir)`` / ``load_state(dir)`` plus an
``epoch.json`` it writes itself (training/training.py:218-250).  ``accelerate`` is a third-party dependency that is
absent from /root/reference (unlisted in requirements.txt, version unpinned), so the directory layout is restated here
from its published behaviour and **parity of the format is unpinned** (no reference test or fixture holds a
checkpoint); what IS pinned, by tests/test_checkpoint_cpu.py, is that the optimizer file loads into a real
``torch.optim.AdamW`` built exactly like training.py:66-82 and comes back bit-identical:

    <dir>/model.safetensors      unwrapped model.state_dict() (reference key names; ``pytorch_model.bin`` is also read)
    <dir>/optimizer.bin          torch.save(optimizer.state_dict()) of AdamW with the two parameter groups of
                                 training.py:73-82: group 0 = gains / biases / logit_scale (no decay), group 1 = the rest
    <dir>/scheduler.bin          torch.save(scheduler.state_dict()); ``last_epoch`` carries the step count
    <dir>/random_states_<rank>.pkl
    <dir>/epoch.json             {"epoch": e, "step": s}   (training.py:223)

The fused optimizer keeps its moments in two flat fp32 buffers (params.ParamStore layout: padded row pitches, backward
order); the functions below translate between that layout and torch's per-parameter state.
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Dict, Iterable, List, Tuple

import torch

from .params import ParamStore, no_decay

MODEL_SAFE, MODEL_BIN, OPT_BIN, SCHED_BIN, EPOCH_JSON = ("model.safetensors", "pytorch_model.bin", "optimizer.bin",
                                                         "scheduler.bin", "epoch.json")


def reference_param_groups(named_shapes: Iterable[Tuple[str, Tuple[int, ...]]]) -> Tuple[List[str], List[str]]:
    """Parameter names of (group 0: no weight decay, group 1: weight decay 0.2) in the order training.py:66-71 builds
    them, i.e. ``named_parameters()`` order filtered by the name rule."""
    named = list(named_shapes)
    g0 = [n for n, s in named if no_decay(n, len(s))]
    g1 = [n for n, s in named if not no_decay(n, len(s))]
    return g0, g1


def _slot_view(store: ParamStore, name: str, flat: torch.Tensor) -> torch.Tensor:
    return store.slots[name].view(flat)


def flat_to_torch_adamw(store: ParamStore, named_shapes, m: torch.Tensor, v: torch.Tensor, t: int, lr: float,
                        betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2) -> dict:
    """``torch.optim.AdamW.state_dict()`` of the reference's optimizer holding the fused optimizer's moments."""
    g0, g1 = reference_param_groups(named_shapes)
    state = {}
    for idx, name in enumerate(g0 + g1):
        if t > 0:
            state[idx] = {"step": torch.tensor(float(t)),
                          "exp_avg": _slot_view(store, name, m).detach().clone().contiguous().cpu(),
                          "exp_avg_sq": _slot_view(store, name, v).detach().clone().contiguous().cpu()}
    common = dict(lr=lr, betas=tuple(betas), eps=eps, amsgrad=False, maximize=False, foreach=None, capturable=False,
                  differentiable=False, fused=None, decoupled_weight_decay=True, initial_lr=lr)
    groups = [dict(common, weight_decay=0.0, params=list(range(len(g0)))),
              dict(common, weight_decay=weight_decay, params=list(range(len(g0), len(g0) + len(g1))))]
    return {"state": state, "param_groups": groups}


def torch_adamw_to_flat(sd: dict, store: ParamStore, named_shapes, m: torch.Tensor, v: torch.Tensor) -> int:
    """Inverse of ``flat_to_torch_adamw``: fills the flat moment buffers (pad elements stay zero), returns the step."""
    named = list(named_shapes)
    g0, g1 = reference_param_groups(named)
    names = g0 + g1
    groups = sd["param_groups"]
    ids = [i for g in groups for i in g["params"]]
    if len(ids) != len(names):
        raise ValueError(f"optimizer state holds {len(ids)} parameters, the model has {len(names)}")
    shapes = dict(named)
    m.zero_()
    v.zero_()
    step = 0
    for pos, idx in enumerate(ids):
        st = sd["state"].get(idx)
        if st is None:
            continue
        name = names[pos]
        if tuple(st["exp_avg"].shape) != tuple(shapes[name]):
            raise ValueError(f"optimizer state {idx} has shape {tuple(st['exp_avg'].shape)}, parameter {name} "
                             f"has {tuple(shapes[name])}")
        _slot_view(store, name, m).copy_(st["exp_avg"])
        _slot_view(store, name, v).copy_(st["exp_avg_sq"])
        step = max(step, int(float(st["step"])))
    return step


def _named_shapes(model):
    return [(n, tuple(p.shape)) for n, p in model.named_parameters()]


def save_state(path: str, model, optimizer, sched_step: int, total_steps: int = 0, rank: int = 0,
               safe_serialization: bool = True) -> None:
    """What ``accelerator.save_state(path)`` leaves behind (training.py:220); rank 0 writes the shared files."""
    os.makedirs(path, exist_ok=True)
    if rank == 0:
        sd = {k: v.detach().cpu().contiguous() for k, v in model.state_dict().items()}
        done = False
        if safe_serialization:
            try:
                from safetensors.torch import save_file
                save_file(sd, os.path.join(path, MODEL_SAFE), metadata={"format": "pt"})
                done = True
            except ImportError:
                pass
        if not done:
            torch.save(sd, os.path.join(path, MODEL_BIN))
        torch.save(flat_to_torch_adamw(optimizer.store, _named_shapes(model), optimizer.m, optimizer.v, optimizer.t,
                                       optimizer.lr, optimizer.betas, optimizer.eps, optimizer.weight_decay),
                   os.path.join(path, OPT_BIN))
        torch.save(scheduler_state_dict(sched_step, total_steps, optimizer.lr), os.path.join(path, SCHED_BIN))
    # accelerate writes this file with torch.save (not pickle.dump) and these keys (checkpointing.save_accelerator_state)
    import random as _random
    import numpy as _np
    states = {"step": int(sched_step), "random_state": _random.getstate(), "numpy_random_seed": _np.random.get_state(),
              "torch_manual_seed": torch.get_rng_state()}
    if torch.cuda.is_available():
        states["torch_cuda_manual_seed"] = torch.cuda.get_rng_state_all()
    torch.save(states, os.path.join(path, f"random_states_{rank}.pkl"))


def scheduler_state_dict(sched_step: int, total_steps: int, max_lr: float = 5e-4, warmup_steps: int = 2):
    """Key set of ``CosineAnnealingWarmupRestarts.state_dict()`` (= the scheduler's __dict__ minus the optimizer) for the
    configuration of training.py:83-89 (cycle_mult 1, gamma 1, min_lr = max_lr / 100) after ``sched_step`` steps.
    The package is absent here (unpinned pip dependency), so the key set is restated from its published source."""
    from .optim import cosine_warmup_lr
    first = max(int(total_steps), warmup_steps + 1)
    min_lr = max_lr / 100
    lr = cosine_warmup_lr(int(sched_step), first, max_lr, min_lr, warmup_steps)
    return {"first_cycle_steps": first, "cycle_mult": 1.0, "base_max_lr": max_lr, "max_lr": max_lr, "min_lr": min_lr,
            "warmup_steps": warmup_steps, "gamma": 1.0, "cur_cycle_steps": first, "cycle": int(sched_step) // first,
            "step_in_cycle": int(sched_step) % first, "base_lrs": [min_lr, min_lr], "last_epoch": int(sched_step),
            "_step_count": int(sched_step) + 1, "verbose": False, "_get_lr_called_within_step": False,
            "_last_lr": [lr, lr]}


def load_state(path: str, model, optimizer) -> int:
    """``accelerator.load_state(path)`` (training.py:243): weights, optimizer moments; returns the scheduler step.
    Raises (FileNotFoundError, ValueError, ...) when the directory is not a checkpoint - the reference's caller
    catches everything and starts from scratch (training.py:245-248)."""
    safe, plain = os.path.join(path, MODEL_SAFE), os.path.join(path, MODEL_BIN)
    if os.path.exists(safe):
        from safetensors.torch import load_file
        sd = load_file(safe)
    else:
        sd = torch.load(plain, map_location="cpu")
    sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}   # DDP-wrapped saves
    model.load_state_dict(sd)
    osd = torch.load(os.path.join(path, OPT_BIN), map_location="cpu", weights_only=False)
    optimizer.t = torch_adamw_to_flat(osd, optimizer.store, _named_shapes(model), optimizer.m, optimizer.v)
    if hasattr(optimizer, "sync_device_counters"):
        optimizer.sync_device_counters(t=optimizer.t)          # the Adam step count also lives on the device
    sched = torch.load(os.path.join(path, SCHED_BIN), map_location="cpu", weights_only=False)
    return int(sched.get("last_epoch", 0))


def write_epoch_json(path: str, epoch: int, step: int) -> None:
    with open(os.path.join(path, EPOCH_JSON), "w") as f:
        json.dump({"epoch": epoch, "step": step}, f)                                  # training.py:223


def read_epoch_json(path: str) -> Dict[str, int]:
    with open(os.path.join(path, EPOCH_JSON)) as f:
        return json.load(f)                                                           # training.py:244
