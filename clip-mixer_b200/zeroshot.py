"""Zero-shot scoring shape of the reference's ImageNet validator (training/clip/validation.py:119-139,142-179;
BASELINE.json configs[4]): per class, encode the prompt templates, L2-normalise, average, re-normalise
(:125-131) -> classifier W [E, classes]; per image batch, encode, normalise, ``100 * f @ W`` (:157-162), top-k
(:136-139).  The encoders are the training-path kernels in inference mode (no activation saving); the small
scoring GEMM runs on the fp32 FFMA engine."""
from __future__ import annotations

from typing import Sequence

import torch

from . import ops
from .ops import MAJOR_K, MAJOR_MN


@torch.no_grad()
def zeroshot_classifier(model, class_token_batches: Sequence[torch.Tensor]) -> torch.Tensor:
    """class_token_batches[c]: int tensor [templates, context_length] of the prompts of class c
    (validation.py:123-124 tokenises 80 templates per class).  Returns W [embed_dim, num_classes] fp32."""
    cols = []
    for tokens in class_token_batches:
        e = model.encode_text(tokens)                                  # :125-128
        e = e / e.norm(dim=-1, keepdim=True)                           # :129
        e = e.mean(dim=0)                                              # :130
        cols.append(e / e.norm())                                      # :131
    return torch.stack(cols, dim=1).contiguous()                       # :133


@torch.no_grad()
def zeroshot_logits(model, images: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """100 * normalise(encode_image(images)) @ W   (validation.py:157-162) -> [batch, classes] fp32."""
    f = model.encode_image(images)
    f = (f / f.norm(dim=-1, keepdim=True)).contiguous()
    B, E = f.shape
    C = W.shape[1]
    out = torch.empty(B, C, device=f.device, dtype=torch.float32)
    # logits[b, c] = sum_e f[b, e] W[e, c]: A = f (K-major), B[n=c][k=e] = W[e, c] (MN-major)
    ops.gemm("simt", B, C, E, 1, f, MAJOR_K, E, 0, W, MAJOR_MN, C, 0, out, C, 0)
    return out.mul_(100.0)


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1, 5)):
    """validation.py:136-139."""
    pred = output.topk(max(topk), 1, True, True)[1].t()
    correct = pred.eq(target.view(1, -1).expand_as(pred))
    return [float(correct[:k].reshape(-1).float().sum().item()) for k in topk]
