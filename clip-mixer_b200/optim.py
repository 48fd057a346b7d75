"""Optimizer side of the training step: the reference's AdamW with its two
parameter groups, clip_grad_norm_(.., 20) and the cosine-with-warm-up schedule, as ONE fused pass over
the model's flat parameter / gradient buffers (training/training.py:66-89,173-186).

``CosineAnnealingWarmupRestarts`` is a third-party dependency of the reference that is absent from
/root/reference (pip package ``cosine_annealing_warmup``, version unpinned, training.py:10,83-89);
its published algorithm is restated in ``cosine_warmup_lr``.
"""
from __future__ import annotations

import math

import torch

from . import ops
from ._lib import MixerClipError


def cosine_warmup_lr(step: int, first_cycle_steps: int, max_lr: float, min_lr: float, warmup_steps: int,
                     cycle_mult: float = 1.0, gamma: float = 1.0) -> float:
    """Learning rate after ``step`` scheduler steps (step 0 = the value at construction)."""
    if first_cycle_steps <= warmup_steps:
        raise ValueError("first_cycle_steps must exceed warmup_steps")
    cycle, cur_steps, in_cycle = 0, first_cycle_steps, step
    if cycle_mult == 1.0:
        cycle = step // first_cycle_steps
        in_cycle = step - cycle * first_cycle_steps
    else:
        while in_cycle >= cur_steps:
            in_cycle -= cur_steps
            cur_steps = int((cur_steps - warmup_steps) * cycle_mult) + warmup_steps
            cycle += 1
    peak = max_lr * (gamma ** cycle)
    if in_cycle < warmup_steps:
        return (peak - min_lr) * in_cycle / warmup_steps + min_lr
    return min_lr + (peak - min_lr) * (1 + math.cos(math.pi * (in_cycle - warmup_steps) / (cur_steps - warmup_steps))) / 2


class FusedAdamW:
    """AdamW over ``model._store`` (flat fp32 params / grads): one mc_sumsq + one mc_adamw launch per step.

    Defaults are the reference's (training.py:73-82): lr 5e-4, betas (0.9, 0.98), eps 1e-6, weight decay 0.2 on
    the tensors its name filter keeps (ParamStore.decay_flags), gradient clipping at 20 (training.py:181).
    The kernel also refreshes the bf16 operand mirror, so the next forward needs no cast pass.
    """

    def __init__(self, model, lr=5e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2, max_grad_norm=20.0):
        store = model._require_store()
        self.model, self.store = model, store
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        dev = store.device
        self.m = torch.zeros(store.total, device=dev)
        self.v = torch.zeros(store.total, device=dev)
        self.sumsq = torch.zeros(1, device=dev)
        # Per-step scalars {lr, 1-b1^t, 1-b2^t} live on the DEVICE and are produced by mc_sched_step from the device
        # counters `state` = {Adam step count t, scheduler step s}, inside the (graph-capturable) step.  Round 1 staged
        # them through ONE pinned host slot per step: with no host sync in the loop the host ran ahead of the GPU and
        # step k+1 could overwrite the slot before step k's async copy had read it.
        self.hyper = torch.zeros(3, device=dev)
        self.state = torch.zeros(2, device=dev, dtype=torch.int64)
        self.t = 0                         # host mirror of state[0] (checkpoints)
        self.schedule = None               # (first_cycle_steps, max_lr, min_lr, warmup_steps) or None = fixed self.lr

    def set_schedule(self, first_cycle_steps: int, max_lr: float, min_lr: float, warmup_steps: int):
        """Cosine-with-warm-up learning rate computed on the device from the scheduler step (training.py:83-89)."""
        if first_cycle_steps <= warmup_steps:
            raise ValueError("first_cycle_steps must exceed warmup_steps")
        self.schedule = (int(first_cycle_steps), float(max_lr), float(min_lr), int(warmup_steps))

    def sync_device_counters(self, t: int = None, sched_step: int = None):
        """Write the host view of {t, s} to the device (after load_state_dict / capture-time tuning replays)."""
        if t is not None:
            self.t = int(t)
        cur = self.state.tolist()
        self.state.copy_(torch.tensor([self.t, cur[1] if sched_step is None else int(sched_step)], dtype=torch.int64))

    def launch_scalars(self, fixed_lr: float = None):
        """Device side, graph-capturable: advance {t, s} and write hyper for this step."""
        if fixed_lr is None and self.schedule is not None:
            f, mx, mn, w = self.schedule
            ops.sched_step(self.state, self.hyper, f, mx, mn, w, self.betas[0], self.betas[1], -1.0)
        else:
            ops.sched_step(self.state, self.hyper, 1, 0.0, 0.0, 0, self.betas[0], self.betas[1],
                           self.lr if fixed_lr is None else fixed_lr)

    def launch(self, grad_mul: float = 1.0):
        """Device side (graph-capturable): grad norm + AdamW + bf16 mirror refresh.  hyper must have been written for
        this step by launch_scalars()."""
        st = self.store
        if st.flat_g is None:
            raise MixerClipError("FusedAdamW.step before any backward")
        mirror = None
        if self.model._precision.act == torch.bfloat16:
            if st.flat_w16 is None:
                st.refresh_mirror(force=True)
            mirror = st.flat_w16
        self.sumsq.zero_()
        ops.sumsq(st.flat_g, st.total, self.sumsq)
        ops.adamw(st.flat_p, st.flat_g, self.m, self.v, mirror, st.decay_flags, st.total, self.sumsq, self.hyper,
                  grad_mul, self.max_grad_norm, self.betas[0], self.betas[1], self.eps, self.weight_decay)
        if mirror is not None:
            st.mirror_is_current()

    def step(self, lr: float = None, grad_mul: float = 1.0):
        self.launch_scalars(self.lr if lr is None else lr)
        self.t += 1
        self.launch(grad_mul)

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm of the last step (device scalar; reading it synchronises)."""
        return self.sumsq.sqrt()

    def zero_grad(self):
        if self.store.flat_g is not None:
            self.store.flat_g.zero_()

    def state_dict(self):
        return {"m": self.m, "v": self.v, "t": self.t}

    def load_state_dict(self, sd):
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.sync_device_counters(t=int(sd["t"]))
