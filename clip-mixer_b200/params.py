"""Flat parameter / gradient storage for the Mixer-CLIP model.

All parameters live in ONE fp32 buffer (and their gradients in a second one, a bf16 operand mirror
in a third), laid out in the order the backward pass finishes them so that gradient all-reduce
buckets are contiguous slices (training/training.py:93,170: DDP bucketed all-reduce) and the fused
optimizer is one launch (training/training.py:73-82,185).

Every 2-D parameter is stored with its row pitch rounded up to 8 elements: the bf16 mirror then
satisfies TMA's 16-byte global-stride rule for the awkward token-mixing shapes (50, 77, 197 tokens:
SURVEY 7.3-1) without any per-step re-packing kernel.  The nn.Parameter the user sees is the
``[:, :cols]`` view, so state-dict shapes and key names are exactly the reference's.
Pad elements are zero and stay zero under AdamW (zero gradient, zero moments).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

CHUNK = 64  # every tensor starts on a 64-element boundary (256 B fp32 / 128 B bf16)


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def no_decay(name: str, ndim: int) -> bool:
    """The reference's weight-decay filter, training/training.py:66."""
    return ndim < 2 or "bn" in name or "ln" in name or "bias" in name or "logit_scale" in name


class Slot:
    __slots__ = ("name", "shape", "rows", "cols", "ld", "offset", "span", "decay")

    def __init__(self, name, shape, offset):
        self.name, self.shape, self.offset = name, tuple(shape), offset
        if len(shape) >= 2:
            self.rows = shape[0]
            self.cols = int(math.prod(shape[1:]))
            self.ld = _round_up(self.cols, 8)
            if len(shape) > 2 and self.ld != self.cols:
                raise ValueError(f"{name}: shape {tuple(shape)} needs a row pitch that is a multiple of 8 "
                                 f"elements (unsupported patch size)")
        else:
            self.rows, self.cols = 1, int(math.prod(shape)) if len(shape) else 1
            self.ld = self.cols
        self.span = _round_up(self.rows * self.ld, CHUNK)
        self.decay = not no_decay(name, len(shape))

    def view(self, flat: torch.Tensor) -> torch.Tensor:
        t = flat[self.offset:self.offset + self.rows * self.ld]
        if len(self.shape) >= 2:
            return t.view(self.rows, self.ld)[:, :self.cols].view(self.shape) if self.ld == self.cols \
                else t.view(self.rows, self.ld)[:, :self.cols]
        return t.view(self.shape)


class ParamStore:
    """Owns the flat buffers and hands out views.  ``order`` is the backward completion order."""

    def __init__(self, shapes: Dict[str, Tuple[int, ...]], order: List[str], device, buckets: List[List[str]]):
        assert sorted(order) == sorted(shapes), "order must name every parameter exactly once"
        self.slots: Dict[str, Slot] = {}
        off = 0
        for name in order:
            s = Slot(name, shapes[name], off)
            self.slots[name] = s
            off += s.span
        self.total = off
        self.order = order
        self.device = torch.device(device)
        self.flat_p = torch.zeros(self.total, device=self.device, dtype=torch.float32)
        self.flat_g = None      # allocated on first backward
        self.flat_w16 = None    # bf16 operand mirror, allocated on first tensor-core forward
        self.w16_version = None
        flags = torch.zeros(self.total // CHUNK, dtype=torch.uint8)
        for s in self.slots.values():
            if s.decay:
                flags[s.offset // CHUNK:(s.offset + s.span) // CHUNK] = 1
        self.decay_flags = flags.to(self.device)
        # contiguous [begin, end) element ranges, one per all-reduce bucket, in backward order
        self.bucket_ranges = []
        for names in buckets:
            b = min(self.slots[n].offset for n in names)
            e = max(self.slots[n].offset + self.slots[n].span for n in names)
            self.bucket_ranges.append((b, e))

    # ---- views -------------------------------------------------------------------------------
    def param_view(self, name):
        return self.slots[name].view(self.flat_p)

    def ensure_grads(self):
        if self.flat_g is None:
            self.flat_g = torch.zeros(self.total, device=self.device, dtype=torch.float32)
        return self.flat_g

    def grad_view(self, name):
        return self.slots[name].view(self.ensure_grads())

    def grad2d(self, name):
        """(tensor starting at the slot, ld) for kernels that write the gradient with its pitch."""
        s = self.slots[name]
        return self.ensure_grads()[s.offset:s.offset + s.rows * s.ld], s.ld

    def weight_operand(self, name, act_dtype):
        """(tensor, ld) of the GEMM operand copy of a 2-D weight: the fp32 parameter itself for the
        SIMT engine, its slice of the bf16 mirror for the tensor-core engine."""
        s = self.slots[name]
        src = self.flat_p if act_dtype == torch.float32 else self.flat_w16
        return src[s.offset:s.offset + s.rows * s.ld], s.ld

    # ---- bf16 mirror ----------------------------------------------------------------------------
    def refresh_mirror(self, force=False):
        from . import ops
        if self.flat_w16 is None:
            self.flat_w16 = torch.empty(self.total, device=self.device, dtype=torch.bfloat16)
            force = True
        ver = self.flat_p._version
        if force or ver != self.w16_version:
            ops.cast_pad(self.flat_p, 1, self.total, self.total, self.flat_w16, self.total)
            self.w16_version = self.flat_p._version

    def mirror_is_current(self):
        """Called by the fused optimizer, which writes the mirror itself."""
        self.w16_version = self.flat_p._version
