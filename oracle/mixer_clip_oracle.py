"""TEST INFRASTRUCTURE ONLY -- functional CPU restatement of the Mixer-CLIP hot path.

This is the parity oracle (see ``oracle/__init__.py`` for who may import it).  It is a
from-scratch functional restatement, over a plain ``dict`` of tensors keyed like the
reference's ``state_dict``, of exactly the arithmetic the reference performs on the path
named by BASELINE.json's north_star:

* LayerNorm / QuickGELU ............. ``training/clip/model.py:166-177``
* MixerBlock (token + channel MLP) ... ``training/clip/model.py:201-222``
* image tower (mixer branch) ......... ``training/clip/model.py:271-290``
* text tower ......................... ``training/clip/model.py:413-426``
* CLIP.forward (normalise, exp) ...... ``training/clip/model.py:428-442``
* contrastive loss with detached
  gathered features .................. ``training/training.py:158-168``

Pinned against the real reference module by ``oracle/make_golden.py`` (fixtures in
``tests/golden/``).  Works in any floating dtype: pass an fp64 state dict for a "truth"
run (the reference's LayerNorm forces fp32, ``model.py:169-172``; here LN computes in
the dtype it is given, which for fp32 inputs is the same thing).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

SOT_ID = 49406  # clip.py:227 via simple_tokenizer: <|startoftext|>
EOT_ID = 49407  # <|endoftext|>, the max id: encode_text finds it with argmax (model.py:424)
LN_EPS = 1e-5   # nn.LayerNorm default used by model.py:166
GELU_A = 1.702  # model.py:177
IMAGE_MEAN = (0.48145466, 0.4578275, 0.40821073)   # training.py:115
IMAGE_STD = (0.26862954, 0.26130258, 0.27577711)   # training.py:115


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
def make_config(embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size,
                context_length, vocab_size, transformer_width, transformer_layers) -> dict:
    """Same quantities as the CLIP ctor arguments that matter in mixer mode (model.py:294-309)."""
    grid = image_resolution // vision_patch_size
    return dict(embed_dim=embed_dim, image_resolution=image_resolution, vision_layers=vision_layers,
                vision_width=vision_width, vision_patch_size=vision_patch_size, grid=grid,
                image_tokens=grid * grid + 1, context_length=context_length, vocab_size=vocab_size,
                transformer_width=transformer_width, transformer_layers=transformer_layers)


CONFIGS = {
    # tiny: full-tensor golden fixture (tests/golden/tiny.pt)
    "tiny": make_config(32, 64, 2, 64, 16, 12, 100, 48, 2),
    # odd: awkward sizes (P=10 / ctx 9, widths 40/24) for edge-case parity
    "odd": make_config(16, 48, 1, 40, 16, 9, 64, 24, 1),
    # S2: BASELINE config-1 shapes ("small": width 512/512, patch 32, 77 tokens) with 2+2 layers
    "S2": make_config(512, 224, 2, 512, 32, 77, 49408, 512, 2),
    # S: BASELINE.json configs[0]
    "S": make_config(512, 224, 12, 512, 32, 77, 49408, 512, 12),
    # B32: training/training.py:275-287 (BASELINE.json configs[1])
    "B32": make_config(512, 224, 12, 768, 32, 77, 49408, 512, 12),
    # B16: B32 with patch 16 (BASELINE.json configs[3])
    "B16": make_config(512, 224, 12, 768, 16, 77, 49408, 512, 12),
}


def infer_config(sd: Dict[str, torch.Tensor]) -> dict:
    """Recover the config from a mixer state dict (what a Mixer-aware build_model must do;
    the reference's build_model, model.py:469-513, only understands transformer keys)."""
    conv = sd["visual.conv1.weight"]
    vision_width, patch = conv.shape[0], conv.shape[-1]
    p_img = sd["visual.transformer.mixBlocks.0.token_mix_seq.lin1.weight"].shape[1]
    grid = int(round(math.sqrt(p_img - 1)))
    vl = len({k.split(".")[3] for k in sd if k.startswith("visual.transformer.mixBlocks.")})
    tl = len({k.split(".")[2] for k in sd if k.startswith("transformer.mixBlocks.")})
    ctx = sd["transformer.mixBlocks.0.token_mix_seq.lin1.weight"].shape[1]
    return make_config(sd["text_projection"].shape[1], patch * grid, vl, vision_width, patch, ctx,
                       sd["token_embedding.weight"].shape[0], sd["ln_final.weight"].shape[0], tl)


def param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    """Every state-dict key of the mixer-mode CLIP and its shape (300 tensors for 12+12 layers)."""
    D, P, E = cfg["vision_width"], cfg["image_tokens"], cfg["embed_dim"]
    W, C = cfg["transformer_width"], cfg["context_length"]
    p = cfg["vision_patch_size"]
    shapes: Dict[str, Tuple[int, ...]] = {
        "text_projection": (W, E), "logit_scale": (),
        "visual.class_embedding": (D,), "visual.proj": (D, E),
        "visual.conv1.weight": (D, 3, p, p),
        "visual.ln_pre.weight": (D,), "visual.ln_pre.bias": (D,),
    }

    def block(prefix, dim, tok):
        shapes[f"{prefix}.layerNorm1.weight"] = (dim,)
        shapes[f"{prefix}.layerNorm1.bias"] = (dim,)
        shapes[f"{prefix}.token_mix_seq.lin1.weight"] = (4 * tok, tok)
        shapes[f"{prefix}.token_mix_seq.lin1.bias"] = (4 * tok,)
        shapes[f"{prefix}.token_mix_seq.lin2.weight"] = (tok, 4 * tok)
        shapes[f"{prefix}.token_mix_seq.lin2.bias"] = (tok,)
        shapes[f"{prefix}.layerNorm2.weight"] = (dim,)
        shapes[f"{prefix}.layerNorm2.bias"] = (dim,)
        shapes[f"{prefix}.channel_mix_seq.lin3.weight"] = (4 * dim, dim)
        shapes[f"{prefix}.channel_mix_seq.lin3.bias"] = (4 * dim,)
        shapes[f"{prefix}.channel_mix_seq.lin4.weight"] = (dim, 4 * dim)
        shapes[f"{prefix}.channel_mix_seq.lin4.bias"] = (dim,)

    for i in range(cfg["vision_layers"]):
        block(f"visual.transformer.mixBlocks.{i}", D, P)
    shapes["visual.ln_post.weight"] = (D,)
    shapes["visual.ln_post.bias"] = (D,)
    for i in range(cfg["transformer_layers"]):
        block(f"transformer.mixBlocks.{i}", W, C)
    shapes["token_embedding.weight"] = (cfg["vocab_size"], W)
    shapes["ln_final.weight"] = (W,)
    shapes["ln_final.bias"] = (W,)
    return shapes


def seeded_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic random weights, bit-identical on every machine (numpy PCG64).

    Not the reference's init RNG (model.py:362-396 is never re-implemented);
    scales are merely chosen in the same range so activations look like a fresh model.
    LayerNorm gains/biases are perturbed away from 1/0 so that their gradients and the
    beta/gamma paths are actually exercised by parity tests.
    """
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if name == "logit_scale":
            arr = np.asarray(math.log(1 / 0.07))
        elif name.endswith("layerNorm1.weight") or name.endswith("layerNorm2.weight") or \
                name in ("visual.ln_pre.weight", "visual.ln_post.weight", "ln_final.weight"):
            arr = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith(".bias"):
            arr = 0.05 * rng.standard_normal(shape)
        elif name == "token_embedding.weight":
            arr = 0.02 * rng.standard_normal(shape)
        else:
            fan_in = shape[-1] if len(shape) == 2 else int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            if name in ("visual.proj", "text_projection"):
                fan_in = shape[0]
            arr = rng.standard_normal(shape) / math.sqrt(fan_in)
        sd[name] = torch.from_numpy(np.asarray(arr, dtype=np.float64)).to(dtype)
    return sd


def synthetic_batch(cfg: dict, batch: int, seed: int = 1, dtype=torch.float32):
    """Parity inputs of SURVEY 8-d: images ~ N(0,1) in the post-normalisation domain; tokens with
    SOT first, one EOT at a random position, zeros after it (unique arg-max as model.py:424 needs)."""
    rng = np.random.default_rng(seed)
    R = cfg["image_resolution"]
    images = torch.from_numpy(rng.standard_normal((batch, 3, R, R))).to(dtype)
    C, V = cfg["context_length"], cfg["vocab_size"]
    eot, sot = V - 1, V - 2
    text = rng.integers(1, max(2, V - 2), size=(batch, C))
    text[:, 0] = sot
    pos = rng.integers(1, C, size=(batch,))
    for b in range(batch):
        text[b, pos[b]] = eot
        text[b, pos[b] + 1:] = 0
    return images, torch.from_numpy(text.astype(np.int64))


# --------------------------------------------------------------------------------------
# forward math
# --------------------------------------------------------------------------------------
def layer_norm(x, weight, bias):
    """model.py:166-172 (biased variance, eps inside the sqrt, affine)."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc * torch.rsqrt(var + LN_EPS) * weight + bias


def quick_gelu(z):
    """model.py:175-177."""
    return z * torch.sigmoid(GELU_A * z)


def mixer_block(x, sd, prefix):
    """model.py:215-222.  x: [B, P, D].  Token-mixing is applied to the transposed [D, P]
    matrix of every sample, channel-mixing to every row."""
    g = lambda k: sd[f"{prefix}.{k}"]
    u = layer_norm(x, g("layerNorm1.weight"), g("layerNorm1.bias"))
    # token mix: Z1[b, h, d] = sum_p W1[h, p] U[b, p, d] + b1[h]          (model.py:220-222)
    z1 = torch.einsum("hp,bpd->bhd", g("token_mix_seq.lin1.weight"), u) + g("token_mix_seq.lin1.bias")[None, :, None]
    h1 = quick_gelu(z1)
    y = x + torch.einsum("ph,bhd->bpd", g("token_mix_seq.lin2.weight"), h1) + g("token_mix_seq.lin2.bias")[None, :, None]
    v = layer_norm(y, g("layerNorm2.weight"), g("layerNorm2.bias"))
    z2 = v @ g("channel_mix_seq.lin3.weight").t() + g("channel_mix_seq.lin3.bias")
    h2 = quick_gelu(z2)
    return y + h2 @ g("channel_mix_seq.lin4.weight").t() + g("channel_mix_seq.lin4.bias")


def patchify(image, patch):
    """Non-overlapping patches as rows: [B,3,R,R] -> [B, g*g, 3*patch*patch], the im2col of the
    stride==kernel convolution at model.py:258,272 (row order gy*g+gx, column order c,py,px)."""
    B, C, R, _ = image.shape
    g = R // patch
    x = image.reshape(B, C, g, patch, g, patch).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(B, g * g, C * patch * patch)


def encode_image(sd, image):
    """model.py:271-290 in mixer mode (no positional embedding, class token kept)."""
    w = sd["visual.conv1.weight"]
    D, patch = w.shape[0], w.shape[-1]
    x = patchify(image.to(w.dtype), patch) @ w.reshape(D, -1).t()          # [B, g*g, D]
    cls = sd["visual.class_embedding"].expand(x.shape[0], 1, D)
    x = torch.cat([cls, x], dim=1)
    x = layer_norm(x, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])
    i = 0
    while f"visual.transformer.mixBlocks.{i}.layerNorm1.weight" in sd:
        x = mixer_block(x, sd, f"visual.transformer.mixBlocks.{i}")
        i += 1
    x = layer_norm(x[:, 0, :], sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])
    return x @ sd["visual.proj"]


def encode_text(sd, text):
    """model.py:413-426 in mixer mode (no positional embedding, no mask)."""
    x = sd["token_embedding.weight"][text]                                 # [B, C, W]
    i = 0
    while f"transformer.mixBlocks.{i}.layerNorm1.weight" in sd:
        x = mixer_block(x, sd, f"transformer.mixBlocks.{i}")
        i += 1
    eot = text.argmax(dim=-1)
    x = x[torch.arange(x.shape[0]), eot]                                   # LN is row-wise: EOT row only
    x = layer_norm(x, sd["ln_final.weight"], sd["ln_final.bias"])
    return x @ sd["text_projection"]


def clip_forward(sd, image, text):
    """model.py:428-442: returns (normalised image feats, normalised text feats, exp(logit_scale))."""
    fi = encode_image(sd, image)
    ft = encode_text(sd, text)
    ui = fi / fi.norm(dim=1, keepdim=True)
    ut = ft / ft.norm(dim=1, keepdim=True)
    return ui, ut, sd["logit_scale"].exp()


def contrastive_loss(ui, ut, scale, ui_all=None, ut_all=None, rank: int = 0):
    """training.py:158-168.  The gathered features are DETACHED; labels are offset by rank."""
    ui_all = ui.detach() if ui_all is None else ui_all.detach()
    ut_all = ut.detach() if ut_all is None else ut_all.detach()
    logits_per_text = scale * ut @ ui_all.t()
    logits_per_image = scale * ui @ ut_all.t()
    n = ui.shape[0]
    gt = torch.arange(n, dtype=torch.long) + rank * n
    ce = torch.nn.functional.cross_entropy
    return (ce(logits_per_image, gt) + ce(logits_per_text, gt)) / 2, logits_per_image, logits_per_text


def loss_and_grads(sd, image, text, world: int = 1):
    """One training step's loss and every parameter gradient.

    ``world`` > 1 emulates W data-parallel ranks on the concatenated global batch: rank r owns
    rows [r*n, (r+1)*n) (split_batches, training.py:64), the loss is the mean over ranks and the
    gradients are DDP-averaged -- which is what this single-process autograd
    computes when every rank's loss is summed and divided by W.
    """
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    ui, ut, s = clip_forward(params, image, text)
    N = ui.shape[0]
    assert N % world == 0
    n = N // world
    total = 0.0
    for r in range(world):
        sl = slice(r * n, (r + 1) * n)
        loss_r, _, _ = contrastive_loss(ui[sl], ut[sl], s, ui, ut, rank=r)
        total = total + loss_r / world
    total.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
    with torch.no_grad():
        _, li, lt = contrastive_loss(ui[:n], ut[:n], s, ui, ut, rank=0)
    return dict(image_features=ui.detach(), text_features=ut.detach(), logit_scale=s.detach(),
                loss=total.detach(), logits_per_image=li, logits_per_text=lt, grads=grads)


def loss_and_grads_chunked(sd, image, text, chunk: int, world: int = 1):
    """Same result as loss_and_grads, computed micro-batch by micro-batch so that full-size batches (BASELINE
    configs[1]: 256 samples) fit in host memory.  Exact, not an approximation: the gathered features are DETACHED
    (training.py:158-159), so the loss of row i depends on row i's feature-with-grad and on all features without
    grad.  Pass 1 computes all features without autograd, pass 2 back-propagates each chunk's rows
    against them with labels offset like training.py:165-167."""
    N = image.shape[0]
    assert N % world == 0 and (N // world) % chunk == 0
    n = N // world
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    with torch.no_grad():
        feats = [clip_forward(params, image[i:i + chunk], text[i:i + chunk]) for i in range(0, N, chunk)]
        ui_all = torch.cat([f[0] for f in feats])
        ut_all = torch.cat([f[1] for f in feats])
    total = 0.0
    for r in range(world):
        for k in range(n // chunk):
            lo = r * n + k * chunk
            ui, ut, s = clip_forward(params, image[lo:lo + chunk], text[lo:lo + chunk])
            # labels of these rows inside rank r's block: arange(chunk) + r*n + k*chunk = "rank" (r*n/chunk + k)
            loss_k, _, _ = contrastive_loss(ui, ut, s, ui_all, ut_all, rank=lo // chunk)
            part = loss_k * (chunk / n) / world
            part.backward()
            total = total + float(part.detach())
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
    with torch.no_grad():
        s = params["logit_scale"].exp()
        _, li, lt = contrastive_loss(ui_all[:n], ut_all[:n], s, ui_all, ut_all, rank=0)
    return dict(image_features=ui_all, text_features=ut_all, logit_scale=s.detach(),
                loss=torch.tensor(total, dtype=ui_all.dtype), logits_per_image=li, logits_per_text=lt, grads=grads)


# --------------------------------------------------------------------------------------
# closed-form head (what the fused CUDA head kernel computes)
# --------------------------------------------------------------------------------------
def head_closed_form(ui, ut, log_scale, ui_all=None, ut_all=None, rank: int = 0):
    """Loss and gradients wrt the *normalised* local features and the log logit-scale, without
    autograd: with P = softmax(s * U_loc @ V_all^T) row-wise and g_i = rank*n+i,
        dU_loc = s/(2n) * (P @ V_all - V_all[g]),   dt = 1/(2n) * sum_i (s*u_i.o_i - A[i, g_i])
    summed over both directions (o_i = sum_j P_ij v_j)."""
    ui_all = ui if ui_all is None else ui_all
    ut_all = ut if ut_all is None else ut_all
    s = math.exp(float(log_scale))
    n = ui.shape[0]
    g = torch.arange(n) + rank * n

    def one(u_loc, v_all):
        A = s * u_loc @ v_all.t()
        lse = torch.logsumexp(A, dim=1)
        P = torch.exp(A - lse[:, None])
        o = P @ v_all
        tgt = A[torch.arange(n), g]
        loss = (lse - tgt).mean()
        du = s / (2 * n) * (o - v_all[g])
        dt = ((s * (u_loc * o).sum(1) - tgt).sum()) / (2 * n)
        return loss, du, dt

    l_i, dui, dt_i = one(ui, ut_all)
    l_t, dut, dt_t = one(ut, ui_all)
    return (l_i + l_t) / 2, dui, dut, dt_i + dt_t


# --------------------------------------------------------------------------------------
# comparison metric (SURVEY 8-c iii)
# --------------------------------------------------------------------------------------
def l2_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2 in fp64 (b is the oracle)."""
    a64, b64 = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    den = b64.norm().item()
    return (a64 - b64).norm().item() / (den if den > 0 else 1.0)


def compare_grads(got: Dict[str, torch.Tensor], ref: Dict[str, torch.Tensor], tol: float):
    """Per-tensor L2-relative error; tensors whose oracle gradient is below 1e-6 of the global
    gradient norm (the analytically-zero token_mix lin2.bias grads) are compared in
    absolute terms against tol * ||global grad|| / sqrt(#tensors).  Returns (worst, failures)."""
    gnorm = math.sqrt(sum(float(v.double().norm()) ** 2 for v in ref.values()))
    floor = 1e-6 * gnorm
    abs_tol = tol * gnorm / math.sqrt(max(1, len(ref)))
    worst, fails = 0.0, []
    for k, r in ref.items():
        g = got[k]
        rn = float(r.double().norm())
        if rn < floor:
            err = float((g.detach().double().cpu() - r.double()).norm())
            ok, shown = err <= abs_tol, err / abs_tol * tol
        else:
            shown = l2_rel(g, r)
            ok = shown <= tol
        worst = max(worst, shown)
        if not ok:
            fails.append((k, shown))
    return worst, fails
