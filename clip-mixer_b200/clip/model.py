"""Drop-in for the Mixer path of the reference's ``training/clip/model.py``.

Same constructor signature, attribute names and state-dict keys as the reference ``CLIP``
(model.py:294-309; 300 tensors, 0 buffers for the 12+12 layer model), same
``encode_image`` / ``encode_text`` / ``forward`` contract -- ``forward`` returns the TRIPLE
``(image_features_normalised, text_features_normalised, logit_scale.exp())`` exactly like
model.py:428-442 -- but the arithmetic runs in libmixerclip's sm_100a kernels.  Only the Mixer
configuration is implemented (``useTransformer=False`` with an int ``vision_layers``); anything
else raises (there is no fallback to a PyTorch implementation and none to CPU).

The sub-modules below are parameter containers that reproduce the reference's module tree (so
``named_parameters()`` drives the reference's weight-decay filter unchanged, training.py:66-71);
the compute is scheduled by ``clip_mixer_b200.engine``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple, Union

import numpy as np
import torch
from torch import nn

from .. import ops
from .._lib import MixerClipError
from ..engine import Precision, TowerRT
from ..params import ParamStore

__all__ = ["CLIP", "build_model", "convert_weights", "LayerNorm", "QuickGELU", "MixerBlock", "Mixer",
           "VisionTransformer", "contrastive_loss"]


# ---------------------------------------------------------------------------------------------
# parameter containers (names follow model.py:166-290)
# ---------------------------------------------------------------------------------------------
class LayerNorm(nn.Module):
    """model.py:166-172 (fp32 statistics always); parameters only, kernel: mc_ln_fwd / mc_ln_bwd."""

    def __init__(self, dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class QuickGELU(nn.Module):
    """model.py:175-177; fused into the GEMM epilogues, kept as a (parameter-free) tree node."""


class _Linear(nn.Module):
    """nn.Linear parameter container with nn.Linear's default initialisation."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1 / math.sqrt(in_features)
        nn.init.uniform_(self.bias, -bound, bound)


class _Conv(nn.Module):
    def __init__(self, width: int, patch: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(width, 3, patch, patch))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))


class _Embedding(nn.Module):
    def __init__(self, vocab: int, width: int):
        super().__init__()
        self.num_embeddings, self.embedding_dim = vocab, width
        self.weight = nn.Parameter(torch.empty(vocab, width))


class MixerBlock(nn.Module):
    """model.py:201-222."""

    def __init__(self, dim: int, num_patch: int):
        super().__init__()
        self.layerNorm1 = LayerNorm(dim)
        self.token_mix_seq = nn.Sequential(OrderedDict([("lin1", _Linear(num_patch, num_patch * 4)),
                                                        ("gelu", QuickGELU()),
                                                        ("lin2", _Linear(num_patch * 4, num_patch))]))
        self.layerNorm2 = LayerNorm(dim)
        self.channel_mix_seq = nn.Sequential(OrderedDict([("lin3", _Linear(dim, dim * 4)), ("gelu", QuickGELU()),
                                                          ("lin4", _Linear(dim * 4, dim))]))


class Mixer(nn.Module):
    """model.py:239-249."""

    def __init__(self, width: int, layers: int, context: int, useGradCheckpointing: bool = False):
        super().__init__()
        self.width, self.layers = width, layers
        self.mixBlocks = nn.Sequential(*[MixerBlock(width, context) for _ in range(layers)])
        self.useGradCheckpointing = useGradCheckpointing


class VisionTransformer(nn.Module):
    """model.py:252-290 in mixer mode (no positional embedding; class token kept)."""

    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int,
                 context_length: int, useTransformer: bool = True):
        super().__init__()
        if useTransformer:
            raise MixerClipError("only the Mixer vision tower (useTransformer=False) is implemented")
        self.useTransformer = False
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.conv1 = _Conv(width, patch_size)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Mixer(width, layers, (input_resolution // patch_size) ** 2 + 1)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self._owner = None

    def forward(self, x: torch.Tensor):
        return self._owner()._encode("image", x, normalise=False)


# ---------------------------------------------------------------------------------------------
# autograd boundary: one Function per tower
# ---------------------------------------------------------------------------------------------
class _TowerFn(torch.autograd.Function):
    """forward: tower(input) -> features (normalised or not); backward: the explicit schedule of
    engine.TowerRT.backward, which writes parameter gradients straight into the flat gradient
    buffer (the gradients returned to autograd for the parameters are None; ``.grad`` is attached
    by the model, see CLIP._attach_grads)."""

    N_FIXED = 5  # model, kind, inp, normalise, save

    @staticmethod
    def forward(ctx, model, kind, inp, normalise, save, *params):
        tower = model._towers[kind]
        model._prepare_weights()
        ws = tower.forward(inp, model._precision, save)
        ctx.model, ctx.kind, ctx.ws, ctx.normalise, ctx.nparams = model, kind, ws, normalise, len(params)
        if save:
            model._fwd_serial[kind] += 1
        ctx.serial = model._fwd_serial[kind]
        ctx.set_materialize_grads(False)
        return (ws.u_feat if normalise else ws.feat).clone()

    @staticmethod
    def backward(ctx, dout):
        model, kind, ws = ctx.model, ctx.kind, ctx.ws
        none = (None,) * (_TowerFn.N_FIXED + ctx.nparams)
        if dout is None:
            return none
        if ctx.serial != model._fwd_serial[kind]:
            raise MixerClipError("backward through a stale forward: the activation arena of this tower was "
                                 "overwritten by a later training forward (one live forward per tower)")
        tower = model._towers[kind]
        model._attach_grads()
        d = dout.to(torch.float32).contiguous()
        hook = model._after_block_hook(kind)
        if ctx.normalise:
            tower.backward(ws, d, model._precision, after_block=hook)
        else:
            tower_backward_feat(tower, ws, d, model._precision, hook)
        return none


def tower_backward_feat(tower: TowerRT, ws, dfeat, prec, after_block):
    """TowerRT.backward with the L2-normalisation step bypassed (inv_norm = 1, u = 0)."""
    saved_u, saved_inv = ws.u_feat, ws.inv_norm
    ws.u_feat = torch.zeros_like(saved_u)
    ws.inv_norm = torch.ones_like(saved_inv)
    try:
        tower.backward(ws, dfeat, prec, after_block=after_block)
    finally:
        ws.u_feat, ws.inv_norm = saved_u, saved_inv


class _HeadFn(torch.autograd.Function):
    """Fused contrastive loss of training.py:158-168 (gathered features detached)."""

    @staticmethod
    def forward(ctx, ui, ut, log_scale, ui_all, ut_all, rank):
        n, E = ui.shape
        N = ui_all.shape[0]
        dev = ui.device
        f = lambda t: t.detach().to(torch.float32).contiguous()
        loss = torch.zeros(1, device=dev)
        dls = torch.zeros(1, device=dev)
        dui, dut = torch.empty(n, E, device=dev), torch.empty(n, E, device=dev)
        ws = torch.empty(ops.head_workspace_bytes(n, N, E) // 4, device=dev)
        ops.head_fwd_bwd(f(ui), f(ut), f(ui_all), f(ut_all), f(log_scale).reshape(1), n, N, E, rank, 1.0, loss, dui,
                         dut, dls, ws)
        ctx.save_for_backward(dui, dut, dls)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        dui, dut, dls = ctx.saved_tensors
        return dui * dloss, dut * dloss, (dls * dloss).reshape(()), None, None, None


def contrastive_loss(image_features, text_features, logit_scale_log, image_features_gathered=None,
                     text_features_gathered=None, rank: int = 0):
    """Fused equivalent of training.py:158-168.  ``logit_scale_log`` is the *parameter* (log of the
    scale); the gathered tensors default to the local ones (single process)."""
    ig = image_features if image_features_gathered is None else image_features_gathered
    tg = text_features if text_features_gathered is None else text_features_gathered
    return _HeadFn.apply(image_features, text_features, logit_scale_log, ig.detach(), tg.detach(), rank)


# ---------------------------------------------------------------------------------------------
# the model
# ---------------------------------------------------------------------------------------------
class CLIP(nn.Module):
    def __init__(self,
                 embed_dim: int,
                 # vision
                 image_resolution: int,
                 vision_layers: Union[Tuple[int, int, int, int], int],
                 vision_width: int,
                 vision_patch_size: int,
                 # text
                 context_length: int,
                 vocab_size: int,
                 transformer_width: int,
                 transformer_heads: int,
                 transformer_layers: int,
                 useTransformer: bool = True,
                 precision: str = "bf16"):
        super().__init__()
        if useTransformer or isinstance(vision_layers, (tuple, list)):
            raise MixerClipError("clip_mixer_b200 implements the Mixer-CLIP path only: construct with "
                                 "useTransformer=False and an int vision_layers (model.py:338)")
        self.context_length = context_length
        self.useTransformer = False
        self.vocab_size = vocab_size
        self.visual = VisionTransformer(image_resolution, vision_patch_size, vision_width, vision_layers,
                                        max(1, vision_width // 64), embed_dim, context_length, useTransformer=False)
        self.transformer = Mixer(width=transformer_width, layers=transformer_layers, context=context_length)
        self.token_embedding = _Embedding(vocab_size, transformer_width)
        self.positional_embedding = None
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        grid = image_resolution // vision_patch_size
        self._cfg = dict(embed_dim=embed_dim, image_resolution=image_resolution, vision_layers=vision_layers,
                         vision_width=vision_width, vision_patch_size=vision_patch_size, grid=grid,
                         image_tokens=grid * grid + 1, context_length=context_length, vocab_size=vocab_size,
                         transformer_width=transformer_width, transformer_layers=transformer_layers)
        if image_resolution % vision_patch_size:
            raise MixerClipError("image_resolution must be a multiple of vision_patch_size")
        self.initialize_parameters()
        self._precision = Precision(precision)
        self._store = None
        self._towers: Dict[str, TowerRT] = {}
        self._tower_params: Dict[str, list] = {}
        self._fwd_serial = {"image": 0, "text": 0}
        self._trusted_mirror = False      # set by the fused trainer, which keeps the bf16 mirror current
        self._dp = None                   # data-parallel hook installer (clip_mixer_b200.dp)
        self._grad_slots = None
        import weakref
        self.visual._owner = weakref.ref(self)
        self._register_state_dict_hook(_clone_state_dict_hook)

    # ---- initialisation: the distributions of model.py:362-396 (not its RNG stream) ----
    def initialize_parameters(self):
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.mixBlocks:
            nn.init.normal_(block.token_mix_seq.lin1.weight, std=fc_std)
            nn.init.normal_(block.token_mix_seq.lin2.weight, std=proj_std)
            nn.init.normal_(block.channel_mix_seq.lin3.weight, std=fc_std)
            nn.init.normal_(block.channel_mix_seq.lin4.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    # ---- flat storage ------------------------------------------------------------------------
    def _flat_order(self):
        L, T = self._cfg["vision_layers"], self._cfg["transformer_layers"]
        blk = ["channel_mix_seq.lin4.weight", "channel_mix_seq.lin4.bias", "channel_mix_seq.lin3.weight",
               "channel_mix_seq.lin3.bias", "layerNorm2.weight", "layerNorm2.bias", "token_mix_seq.lin2.weight",
               "token_mix_seq.lin2.bias", "token_mix_seq.lin1.weight", "token_mix_seq.lin1.bias", "layerNorm1.weight",
               "layerNorm1.bias"]
        # Backward order of the fused step: text tower first (its 101 MB embedding-gradient bucket then
        # overlaps the whole image-tower backward), image tower second, logit_scale last (its gradient may
        # arrive from torch autograd at any time, so it is reduced by DataParallel.finish()).
        buckets, tags = [["text_projection", "ln_final.weight", "ln_final.bias"]], [("text", "top")]
        for i in range(T - 1, -1, -1):
            buckets.append([f"transformer.mixBlocks.{i}.{n}" for n in blk])
            tags.append(("text", i))
        buckets.append(["token_embedding.weight"])
        tags.append(("text", "bottom"))
        buckets.append(["visual.proj", "visual.ln_post.weight", "visual.ln_post.bias"])
        tags.append(("image", "top"))
        for i in range(L - 1, -1, -1):
            buckets.append([f"visual.transformer.mixBlocks.{i}.{n}" for n in blk])
            tags.append(("image", i))
        buckets.append(["visual.ln_pre.weight", "visual.ln_pre.bias", "visual.class_embedding", "visual.conv1.weight"])
        tags.append(("image", "bottom"))
        buckets.append(["logit_scale"])
        tags.append(("head", "final"))
        return buckets, tags

    def _rebuild_store(self):
        """(Re)create the flat buffers on the parameters' device and re-point every Parameter."""
        params = dict(self.named_parameters())
        dev = self.logit_scale.device
        for n, p in params.items():
            if p.dtype != torch.float32:
                raise MixerClipError(f"parameter {n} is {p.dtype}: master weights stay fp32; choose the compute "
                                     "precision with set_precision('bf16'|'fp32') instead of .half()/.bfloat16()")
        for t in self._towers.values():
            t.release()
        self._towers, self._store = {}, None
        if dev.type != "cuda":
            return
        buckets, tags = self._flat_order()
        order = [n for b in buckets for n in b]
        store = ParamStore({n: tuple(p.shape) for n, p in params.items()}, order, dev, buckets)
        with torch.no_grad():
            for n, p in params.items():
                view = store.param_view(n)
                view.copy_(p.data)
                p.data = view
                p.grad = None
        self._store, self._bucket_tags, self._grad_slots = store, tags, None
        self._tower_params = {
            "image": [p for n, p in params.items() if n.startswith("visual.")],
            "text": [p for n, p in params.items() if not n.startswith("visual.") and n != "logit_scale"],
        }
        self._towers = {k: TowerRT(k, self._cfg, params, store) for k in ("image", "text")}

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._rebuild_store()
        return out

    def _require_store(self):
        if self._store is None:
            if self.logit_scale.device.type != "cuda":
                raise MixerClipError("clip_mixer_b200 runs on CUDA (sm_100a) only: move the model with "
                                     ".to('cuda') first. There is no CPU fallback.")
            self._rebuild_store()
        return self._store

    def _prepare_weights(self):
        store = self._require_store()
        if self._precision.act == torch.bfloat16:
            untrusted_training = self.training and torch.is_grad_enabled() and not self._trusted_mirror
            store.refresh_mirror(force=untrusted_training)

    def mark_weights_dirty(self):
        """Call after changing weights through ``.data`` (bypasses version tracking)."""
        if self._store is not None:
            self._store.w16_version = None

    def _attach_grads(self):
        """Point every ``p.grad`` at its slice of the flat gradient buffer.  A ``None`` grad (after
        ``zero_grad(set_to_none=True)``) means that slice must restart from zero."""
        store = self._require_store()
        if store.flat_g is None or self._grad_slots is None:
            store.ensure_grads()
            self._grad_slots = [(p, store.grad_view(n)) for n, p in self.named_parameters()]
            self._grad_slots = [(p, gv, gv.data_ptr()) for p, gv in self._grad_slots]
        with torch.no_grad():
            for p, gv, ptr in self._grad_slots:
                g = p.grad
                if g is None:
                    gv.zero_()
                    p.grad = gv
                elif g.data_ptr() != ptr:
                    gv.copy_(g)
                    p.grad = gv

    def _after_block_hook(self, kind):
        return self._dp.after_block_hook(kind) if self._dp is not None else None

    def set_precision(self, name: str):
        """'bf16' (tcgen05 tensor cores, default) or 'fp32' (SIMT FFMA, the 1e-5 validation mode)."""
        self._precision = Precision(name)
        for t in self._towers.values():
            t.release()
        return self

    # ---- reference surface ----------------------------------------------------------------------
    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype                                  # model.py:406-408

    def _encode(self, kind, inp, normalise):
        self._require_store()
        if not inp.is_cuda:
            raise MixerClipError("inputs must be CUDA tensors (no CPU fallback)")
        params = self._tower_params[kind]
        save = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _TowerFn.apply(self, kind, inp, normalise, save, *params)

    def encode_image(self, image):
        return self._encode("image", image, normalise=False)                    # model.py:410-411

    def encode_text(self, text: torch.Tensor):
        return self._encode("text", text, normalise=False)                      # model.py:413-426

    def forward(self, image, text):
        image_features = self._encode("image", image, normalise=True)           # model.py:430-434
        text_features = self._encode("text", text, normalise=True)
        logit_scale = self.logit_scale.exp()                                    # model.py:437
        return image_features, text_features, logit_scale                       # model.py:442


def _clone_state_dict_hook(module, state_dict, prefix, local_metadata):
    """Parameters are views into one flat buffer; hand out independent contiguous tensors so that
    torch.save does not serialise the whole buffer once per key."""
    for k in list(state_dict.keys()):
        if k.startswith(prefix):
            state_dict[k] = state_dict[k].detach().clone().contiguous()
    return state_dict


def convert_weights(model: nn.Module):
    """model.py:445-466 casts weights to fp16 for inference.  Here the compute precision is a
    property of the engine (bf16 tensor cores over fp32 master weights), so this selects it."""
    if isinstance(model, CLIP):
        model.set_precision("bf16")
    return model


def build_model(state_dict: dict):
    """Mixer-aware counterpart of model.py:469-513 (the reference's version only understands
    transformer checkpoints): infer the configuration from the ``mixBlocks`` keys."""
    if not any(".mixBlocks." in k for k in state_dict):
        raise MixerClipError("state dict has no mixBlocks.* keys: only Mixer-CLIP checkpoints are supported")
    sd = {k: v for k, v in state_dict.items() if k not in ("input_resolution", "context_length", "vocab_size")}
    conv = sd["visual.conv1.weight"]
    vision_width, patch = conv.shape[0], conv.shape[-1]
    p_img = sd["visual.transformer.mixBlocks.0.token_mix_seq.lin1.weight"].shape[1]
    grid = int(round(math.sqrt(p_img - 1)))
    vision_layers = len({k.split(".")[3] for k in sd if k.startswith("visual.transformer.mixBlocks.")})
    transformer_layers = len({k.split(".")[2] for k in sd if k.startswith("transformer.mixBlocks.")})
    context_length = sd["transformer.mixBlocks.0.token_mix_seq.lin1.weight"].shape[1]
    transformer_width = sd["ln_final.weight"].shape[0]
    model = CLIP(sd["text_projection"].shape[1], patch * grid, vision_layers, vision_width, patch, context_length,
                 sd["token_embedding.weight"].shape[0], transformer_width, max(1, transformer_width // 64),
                 transformer_layers, useTransformer=False)
    model.load_state_dict({k: v.float() for k, v in sd.items()})
    return model.eval()
