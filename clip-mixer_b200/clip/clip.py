"""Drop-in for the public API of the reference's ``training/clip/clip.py``:
``available_models`` (:90-92), ``load`` (:95-195), ``tokenize`` (:198-238), ``_transform`` (:80-87).

Differences that follow from the scope:
  * the named models are Mixer configurations constructed with random weights -- the reference
    publishes no Mixer checkpoint and its ``load`` cannot read one (build_model is transformer-only);
    a path to a Mixer state dict (or an accelerate-style checkpoint holding one) loads through the
    Mixer-aware ``build_model``;
  * ``jit=True`` is rejected (TorchScript archives of the OpenAI transformer models are out of scope);
  * ``tokenize`` pads / truncates exactly like the reference (SOT first, EOT last, zero padding, int32
    result) and accepts token id sequences or raw strings.  Raw strings are BPE-encoded by ``clip/bpe.py``
    (pinned against the reference tokenizer, tests/golden/bpe.json); the merge table is a data asset of
    the reference (``bpe_simple_vocab_16e6.txt.gz``), found through ``CLIP_BPE_VOCAB`` or ``baseline/_ref``
    -- without it raw strings raise and token ids must be passed.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Union

import numpy as np
import torch

from .._lib import MixerClipError
from .model import CLIP, build_model

__all__ = ["available_models", "load", "tokenize"]

SOT_TOKEN = 49406
EOT_TOKEN = 49407

# name -> CLIP constructor arguments (training/training.py:275-287 is "Mixer-B/32")
_MODELS = {
    "Mixer-B/32": dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=32,
                       context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
                       transformer_layers=12),
    "Mixer-B/16": dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=16,
                       context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
                       transformer_layers=12),
    "Mixer-S/32": dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=512, vision_patch_size=32,
                       context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
                       transformer_layers=12),
}


def available_models() -> List[str]:
    """Returns the names of available CLIP models (clip.py:90-92)."""
    return list(_MODELS.keys())


def _convert_image_to_rgb(image):
    return np.moveaxis(np.array(image.convert("RGB")), -1, 0)


def _transform(n_px: int):
    """clip.py:80-87: resize (bicubic), centre crop, uint8 CHW array; the /255 + Normalize of
    training.py:115,149 is fused into the patch-embedding operand producer (mc_im2col)."""
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Resize
    return Compose([Resize(n_px, interpolation=InterpolationMode.BICUBIC), CenterCrop(n_px), _convert_image_to_rgb])


def load(name: str, device: Union[str, torch.device] = "cuda" if torch.cuda.is_available() else "cpu",
         jit: bool = False, download_root: str = None):
    """Load a Mixer-CLIP model; returns ``(model, preprocess)`` like clip.py:95-143."""
    if jit:
        raise MixerClipError("jit=True (TorchScript archives) is not supported by clip_mixer_b200")
    if name in _MODELS:
        model = CLIP(**_MODELS[name], useTransformer=False)
    elif os.path.isfile(name):
        obj = torch.load(name, map_location="cpu")
        state_dict = obj.get("state_dict", obj.get("model", obj)) if isinstance(obj, dict) else obj.state_dict()
        state_dict = {k[len("module."):] if k.startswith("module.") else k: v for k, v in state_dict.items()}
        model = build_model(state_dict)
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")   # clip.py:125
    model = model.to(device)
    return model, _transform(model.visual.input_resolution)


def tokenize(texts: Union[str, List[str], Sequence[Sequence[int]], torch.Tensor], context_length: int = 77,
             truncate: bool = False) -> torch.Tensor:
    """clip.py:198-238.  Accepts raw strings (BPE-encoded by clip/bpe.py; needs the reference's merge table, see the
    module docstring) or already-encoded token id sequences (without SOT/EOT).  Result: int32 [len(texts), context_length]."""
    if isinstance(texts, str):
        texts = [texts]
    if isinstance(texts, torch.Tensor):
        texts = texts.tolist()
    encoded = []
    for t in texts:
        if isinstance(t, str):
            from . import bpe
            encoded.append(bpe.encode(t))
        else:
            encoded.append([int(v) for v in t])
    result = torch.zeros(len(encoded), context_length, dtype=torch.int)
    for i, ids in enumerate(encoded):
        tokens = [SOT_TOKEN] + ids + [EOT_TOKEN]
        if len(tokens) > context_length:
            if truncate:
                tokens = tokens[:context_length]
                tokens[-1] = EOT_TOKEN
            else:
                raise RuntimeError(f"Input {texts[i]} is too long for context length {context_length}")  # clip.py:235
        result[i, :len(tokens)] = torch.tensor(tokens)
    return result
