"""Execution engine of the Mixer-CLIP hot path: explicit forward / backward schedules of the two
Mixer towers over libmixerclip kernels (no autograd inside; torch.autograd sees one Function per
tower, see clip/model.py).

Math follows SURVEY.md Appendix A, which restates training/clip/model.py:215-222 (MixerBlock),
:271-290 (image tower), :413-426 (text tower).  Layout decisions:
  * residual stream, LayerNorm statistics, features, every gradient of a parameter: fp32
  * GEMM operands ("act" tensors): bf16 with the tcgen05 engine, fp32 with the SIMT engine
  * activations stay [B, P, D] row-major for BOTH MLPs: the token-mixing GEMMs read that tensor as
    an MN-major operand through the TMA descriptor, no transposed copy is ever made
    (reference: x.permute(0, 2, 1) + contiguous copy, model.py:220-222)
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from . import ops
from ._lib import MixerClipError
from .ops import (ACT_GELU, ACT_GELU_BWD, ACT_NONE, BIAS_M, BIAS_N, MAJOR_K, MAJOR_MN)
from .params import ParamStore


def fused_ln_prologue_enabled() -> bool:
    """MC_TM_FUSE_LN=1: LayerNorm 1 of blocks 1.. runs in the prologue of the fused token-mixing forward kernel, with row
    statistics from the preceding lin4 GEMM's epilogue, instead of as its own kernel.  Off by default: parity-green and
    22 launches fewer, but the store warps that do the normalising are the forward kernel's bottleneck - interleaved A/B
    14.46 / 14.92 ms without against 15.17 / 14.84 ms with (profiles/r2i_fuse_ln_ab.txt)."""
    return os.environ.get("MC_TM_FUSE_LN", "0") == "1"


def fused_token_mix_enabled() -> bool:
    """MC_TOKENMIX=gemm selects the unfused token-mixing schedule (separate engine GEMMs with a materialised
    [B, 4P, D] hidden tensor) for A/B measurements; the default is the fused kernels of csrc/tokenmix.cu."""
    return os.environ.get("MC_TOKENMIX", "fused") != "gemm"


class Precision:
    """bf16: tcgen05 tensor-core GEMMs on bf16 operands (fp32 accumulate); fp32: SIMT FFMA GEMMs."""

    def __init__(self, name: str):
        if name not in ("bf16", "fp32"):
            raise MixerClipError(f"precision must be 'bf16' or 'fp32', got {name!r}")
        self.name = name
        self.engine = "tc" if name == "bf16" else "simt"
        self.act = torch.bfloat16 if name == "bf16" else torch.float32
        # saved pre-activations: fp16 (same bytes as bf16, 8x finer) so that g'(z) in backward adds no
        # rounding noise comparable to an operand rounding
        self.z = torch.float16 if name == "bf16" else torch.float32
        # tensor-core engine: token-mixing backward recomputes Z1 = W1 U + b1 inside the dZ1 GEMM (second operand
        # pair), so the forward pass does not store it; the fp32 validation engine keeps the saved copy
        self.recompute_z1 = name == "bf16"


class TowerWS:
    """Per-(tower, batch) activation arena.  With save=False (inference) the per-layer tensors
    collapse to one slot that every layer reuses."""

    def __init__(self, B, P, D, E, L, act, zdt, device, save, keep_z1=True, fused_tm=False):
        self.B, self.save, self.fused_tm = B, save, fused_tm
        n = L if save else 1
        f32 = dict(device=device, dtype=torch.float32)
        a = dict(device=device, dtype=act)
        self.x = torch.empty((L + 1) if save else 2, B, P, D, **f32)   # block inputs (x[L] = tower output)
        self.y = torch.empty(n, B, P, D, **f32)                       # mid-block residual
        self.u = torch.empty(n, B, P, D, **a)
        self.v = torch.empty(n, B, P, D, **a)
        # fused token mixing keeps the hidden activation on chip (forward) and recomputes it (backward)
        self.z1 = torch.empty(n, B, 4 * P, D, device=device, dtype=zdt) if (save and keep_z1 and not fused_tm) else None
        self.h1 = torch.empty(n, B, 4 * P, D, **a) if not fused_tm else None
        self.z2 = torch.empty(n, B * P, 4 * D, device=device, dtype=zdt) if save else None
        self.h2 = torch.empty(n, B * P, 4 * D, **a)
        self.stats = torch.empty(n, 4, B * P, **f32)                  # mean1, rstd1, mean2, rstd2
        # fused LayerNorm prologue: (sum, sum of squares) over D of every row of a block's input, left behind by the lin4 GEMM
        # that produced it (slot i = input of block i; slot 0 is unused: block 0's input comes from ln_pre / the embedding)
        self.lnsum = torch.zeros((L + 1) if save else 2, B * P, 2, **f32) if fused_tm else None
        self.top_a = torch.empty(B, D, **f32)                         # LN(cls / EOT row), projection operand (fp32)
        self.top_stats = torch.empty(2, B, **f32)
        self.rows = torch.empty(B, device=device, dtype=torch.int32)  # cls / EOT row index per sample
        self.feat = torch.empty(B, E, **f32)
        self.u_feat = torch.empty(B, E, **f32)                        # L2-normalised features
        self.inv_norm = torch.empty(B, **f32)
        self.extra: Dict[str, torch.Tensor] = {}
        # backward scratch (allocated lazily)
        self.bwd = None

    def alloc_bwd(self, P, D, E, act, device):
        if self.bwd is not None:
            return self.bwd
        B = self.B
        f32 = dict(device=device, dtype=torch.float32)
        a = dict(device=device, dtype=act)
        s = {}
        s["dcur"] = torch.empty(B, P, D, **f32)
        s["dcur_a"] = torch.empty(B, P, D, **a) if act != torch.float32 else s["dcur"]
        s["dtmp"] = torch.empty(B, P, D, **f32)
        s["dz2"] = torch.empty(B * P, 4 * D, **a)
        s["dz1"] = torch.empty(B, (1 if self.fused_tm else 4) * P, D, **a)   # fused: only the image tower's operand scratch
        s["dfeat"] = torch.empty(B, E, **f32)
        self.bwd = s
        return s


class TowerRT:
    """One Mixer tower (image or text) bound to its parameters."""

    def __init__(self, kind: str, cfg: dict, params: Dict[str, torch.nn.Parameter], store: ParamStore):
        assert kind in ("image", "text")
        self.kind, self.cfg, self.p, self.store = kind, cfg, params, store
        if kind == "image":
            self.D, self.P, self.L = cfg["vision_width"], cfg["image_tokens"], cfg["vision_layers"]
            self.blk = "visual.transformer.mixBlocks"
        else:
            self.D, self.P, self.L = cfg["transformer_width"], cfg["context_length"], cfg["transformer_layers"]
            self.blk = "transformer.mixBlocks"
        self.E = cfg["embed_dim"]
        if self.D % 8 or self.E % 8:
            raise MixerClipError(f"{kind} tower: width {self.D} and embed_dim {self.E} must be multiples of 8")
        if self.D > 1024 or self.E > 512:
            raise MixerClipError(f"{kind} tower: width {self.D} > 1024 or embed_dim {self.E} > 512 unsupported")
        self._ws: Dict[tuple, TowerWS] = {}

    # ---- helpers --------------------------------------------------------------------------------
    def bp(self, i, name):
        return self.p[f"{self.blk}.{i}.{name}"]

    def wop(self, name, prec):
        return self.store.weight_operand(name, prec.act)

    def workspace(self, B, prec: Precision, save: bool) -> TowerWS:
        key = (B, prec.name, save, self.fused_tm(prec))
        ws = self._ws.get(key)
        if ws is None:
            # keep at most one training and one inference arena per tower
            for k in [k for k in self._ws if k[2] == save]:
                del self._ws[k]
            ws = TowerWS(B, self.P, self.D, self.E, self.L, prec.act, prec.z, self.store.device, save,
                         keep_z1=not prec.recompute_z1, fused_tm=self.fused_tm(prec))
            self._ws[key] = ws
        return ws

    def release(self):
        self._ws.clear()

    def fused_tm(self, prec: Precision) -> bool:
        return prec.engine == "tc" and fused_token_mix_enabled() and ops.token_mix_supported(self.P, self.D)

    def refresh_w1t(self, prec: Precision):
        """W1^T bf16 copies [L, P, ld1t] of the token-mixing lin1 weights for the fused kernels, whose resident weight
        tiles are fetched by TMA (csrc/tokenmix.cu).  Refreshed at the start of every forward pass (the weights change
        once per step): ONE batched launch when the blocks' operands are evenly spaced in the bf16 mirror (they are:
        every block has the same 12 tensors in the same order), else one launch per block."""
        P, L = self.P, self.L
        H = 4 * P
        self.ld1t = (H + 7) // 8 * 8
        if getattr(self, "w1t", None) is None or self.w1t.device != self.store.device:
            self.w1t = torch.zeros(L, P, self.ld1t, device=self.store.device, dtype=torch.bfloat16)
        ops_ = [self.wop(f"{self.blk}.{i}.token_mix_seq.lin1.weight", prec) for i in range(L)]
        ptrs = [w.data_ptr() for w, _ in ops_]
        ld1 = ops_[0][1]
        step = (ptrs[1] - ptrs[0]) if L > 1 else 0
        even = all(ld == ld1 for _, ld in ops_) and all(ptrs[i] - ptrs[0] == i * step for i in range(L)) and step % 2 == 0
        if even and L > 1 and step != 0:
            # block i lives at ptrs[0] + i * step (step < 0: the flat layout follows the backward order, block L-1 first)
            ops.transpose_bf16(ops_[0][0], H, P, ld1, step // 2, self.w1t, self.ld1t, P * self.ld1t, L)
        else:
            for i, (w, ld) in enumerate(ops_):
                ops.transpose_bf16(w, H, P, ld, 0, self.w1t[i], self.ld1t, 0, 1)

    # ---- forward --------------------------------------------------------------------------------
    def forward(self, inp: torch.Tensor, prec: Precision, save: bool) -> TowerWS:
        B = inp.shape[0]
        ws = self.workspace(B, prec, save)
        D, P, L, E, eng = self.D, self.P, self.L, self.E, prec.engine
        cfg = self.cfg
        if self.kind == "image":
            R, patch, g = cfg["image_resolution"], cfg["vision_patch_size"], cfg["grid"]
            if tuple(inp.shape) != (B, 3, R, R):
                raise MixerClipError(f"image must be [B,3,{R},{R}], got {tuple(inp.shape)}")
            Kc = 3 * patch * patch
            key = "patches"
            if key not in ws.extra:
                ws.extra[key] = torch.empty(B * g * g, Kc, device=inp.device, dtype=prec.act)
                ws.extra["xpre"] = torch.empty(B, P, D, device=inp.device, dtype=torch.float32)
                ws.extra["pre_stats"] = torch.empty(2, B * P, device=inp.device, dtype=torch.float32)
                ws.rows.copy_(torch.arange(B, device=inp.device, dtype=torch.int32) * P)
            patches, xpre, pst = ws.extra["patches"], ws.extra["xpre"], ws.extra["pre_stats"]
            img = inp if inp.dtype == torch.uint8 else inp.to(torch.float32)
            ops.im2col(img.contiguous(), B, R, patch, patches)                                # model.py:272
            wc, ldc = self.wop("visual.conv1.weight", prec)
            ops.gemm(eng, B * g * g, D, Kc, 1, patches, MAJOR_K, Kc, 0, wc, MAJOR_K, ldc, 0, xpre, D, 0,
                     row_remap=g * g)                                                        # rows 1.. of each sample
            ops.ln_fwd(xpre, D, self.p["visual.ln_pre.weight"], self.p["visual.ln_pre.bias"], ws.x[0], D, pst[0],
                       pst[1], B * P, D, cls=self.p["visual.class_embedding"], cls_period=P)  # model.py:275-279
        else:
            C, V = cfg["context_length"], cfg["vocab_size"]
            if tuple(inp.shape) != (B, C):
                raise MixerClipError(f"text must be [B,{C}], got {tuple(inp.shape)}")
            text = inp.to(torch.int64).contiguous()
            ws.extra["text"] = text
            ops.embed_fwd(text, self.p["token_embedding.weight"], ws.x[0], B, C, D, V)        # model.py:414
            ops.eot_rows(text, ws.rows, B, C)                                                # model.py:424

        if ws.fused_tm:
            self.refresh_w1t(prec)
            if fused_ln_prologue_enabled():
                ws.lnsum.zero_()
        for i in range(L):
            self._block_fwd(i, ws, prec, save)

        xo = ws.x[L] if save else ws.x[L % 2]
        if self.kind == "image":
            gam, bet, proj = self.p["visual.ln_post.weight"], self.p["visual.ln_post.bias"], "visual.proj"
        else:
            gam, bet, proj = self.p["ln_final.weight"], self.p["ln_final.bias"], "text_projection"
        ops.ln_fwd(xo, D, gam, bet, ws.top_a, D, ws.top_stats[0], ws.top_stats[1], B, D, row_index=ws.rows)
        # The two projections are 0.01 % of the FLOPs but set the error of the features that feed the
        # logits: they always run on the fp32 FFMA engine from the fp32 master weights.
        wp, ldp = self.store.weight_operand(proj, torch.float32)                              # [D, E]
        ops.gemm("simt", B, E, D, 1, ws.top_a, MAJOR_K, D, 0, wp, MAJOR_MN, ldp, 0, ws.feat, E, 0)  # model.py:288,424
        ops.l2norm_fwd(ws.feat, ws.u_feat, ws.inv_norm, B, E)                                 # model.py:433-434
        return ws

    def _block_fwd(self, i, ws: TowerWS, prec: Precision, save: bool):
        B, D, P, eng = ws.B, self.D, self.P, prec.engine
        s = i if save else 0
        x = ws.x[i] if save else ws.x[i % 2]
        xo = ws.x[i + 1] if save else ws.x[(i + 1) % 2]
        y, u, v, h2, st = ws.y[s], ws.u[s], ws.v[s], ws.h2[s], ws.stats[s]
        h1 = ws.h1[s] if ws.h1 is not None else None
        z1 = ws.z1[s] if (save and ws.z1 is not None) else None
        z2 = ws.z2[s] if save else None
        pre = f"{self.blk}.{i}."
        # x + token_mix(LN1(x))                                                          model.py:216,220-222
        # Blocks 1..: LayerNorm 1 runs in the PROLOGUE of the fused token-mixing kernel - the preceding lin4 GEMM left the
        # row sums of x in ws.lnsum (rowstat_out), the kernel normalises the fp32 tile it reads anyway for the residual.
        fuse_ln = ws.fused_tm and i > 0 and fused_ln_prologue_enabled()
        if not fuse_ln:
            ops.ln_fwd(x, D, self.bp(i, "layerNorm1.weight"), self.bp(i, "layerNorm1.bias"), u, D, st[0], st[1], B * P, D)
        w1, ld1 = self.wop(pre + "token_mix_seq.lin1.weight", prec)                       # [4P, P]
        w2, ld2 = self.wop(pre + "token_mix_seq.lin2.weight", prec)                       # [P, 4P]
        if ws.fused_tm:
            # one kernel: both GEMMs, bias, QuickGELU and the residual; the hidden [4P x D] tile never leaves the SM
            ln = None
            if fuse_ln:
                ln = dict(sums=ws.lnsum[i if save else i % 2], gamma=self.bp(i, "layerNorm1.weight"),
                          beta=self.bp(i, "layerNorm1.bias"), u_out=u, mean=st[0], rstd=st[1])
            ops.token_mix_fwd(B, P, D, u, x, y, w1, ld1, self.bp(i, "token_mix_seq.lin1.bias"), w2, ld2,
                              self.bp(i, "token_mix_seq.lin2.bias"), w1t=self.w1t[i], ld1t=self.ld1t, ln=ln)
        else:
            ops.gemm(eng, 4 * P, D, P, B, w1, MAJOR_K, ld1, 0, u, MAJOR_MN, D, P * D, h1, D, 4 * P * D,
                     bias=self.bp(i, "token_mix_seq.lin1.bias"), bias_mode=BIAS_M, zout=z1, ldz=D, z_bs=4 * P * D,
                     act=ACT_GELU)
            ops.gemm(eng, P, D, 4 * P, B, w2, MAJOR_K, ld2, 0, h1, MAJOR_MN, D, 4 * P * D, y, D, P * D,
                     bias=self.bp(i, "token_mix_seq.lin2.bias"), bias_mode=BIAS_M, R=x, ldr=D, r_bs=P * D)
        # y + channel_mix(LN2(y))                                                         model.py:217
        ops.ln_fwd(y, D, self.bp(i, "layerNorm2.weight"), self.bp(i, "layerNorm2.bias"), v, D, st[2], st[3], B * P, D)
        w3, ld3 = self.wop(pre + "channel_mix_seq.lin3.weight", prec)                     # [4D, D]
        ops.gemm(eng, B * P, 4 * D, D, 1, v, MAJOR_K, D, 0, w3, MAJOR_K, ld3, 0, h2, 4 * D, 0,
                 bias=self.bp(i, "channel_mix_seq.lin3.bias"), bias_mode=BIAS_N, zout=z2, ldz=4 * D, act=ACT_GELU)
        w4, ld4 = self.wop(pre + "channel_mix_seq.lin4.weight", prec)                     # [D, 4D]
        nxt = None                                            # row sums of xo for the next block's fused LayerNorm
        if ws.fused_tm and i + 1 < self.L and fused_ln_prologue_enabled():
            nxt = ws.lnsum[(i + 1) if save else (i + 1) % 2]
            if not save:
                nxt.zero_()                                   # inference: two alternating slots, re-zeroed before every use
        ops.gemm(eng, B * P, D, 4 * D, 1, h2, MAJOR_K, 4 * D, 0, w4, MAJOR_K, ld4, 0, xo, D, 0,
                 bias=self.bp(i, "channel_mix_seq.lin4.bias"), bias_mode=BIAS_N, R=y, ldr=D, rowstat_out=nxt)

    # ---- backward -------------------------------------------------------------------------------
    def backward(self, ws: TowerWS, du_feat: torch.Tensor, prec: Precision, after_block=None):
        """du_feat: gradient wrt the L2-normalised features [B, E] fp32.  Accumulates every parameter
        gradient of this tower into the flat gradient buffer.  after_block(tag) is called when a
        group of parameters is complete (tag = "top", block index, "bottom") - the data-parallel
        wrapper launches that bucket's all-reduce from it."""
        for _ in self.backward_iter(ws, du_feat, prec, after_block):
            pass

    def backward_iter(self, ws: TowerWS, du_feat: torch.Tensor, prec: Precision, after_block=None):
        """Generator form of backward(): yields after every parameter group ("top", each block, "bottom") so that the
        caller can interleave the HOST-side enqueue of the two towers' backward passes.  The gradient all-reduces share
        one in-order communication stream: enqueued tower after tower, every image-tower bucket queued behind the LAST
        text-tower bucket (measured on 2 GPUs, profiles/r2c_bench_2gpu.json: image buckets ready at 4.8 .. 12.8 ms all
        started after 13.2 ms); interleaved, the stream order follows the order in which buckets become ready."""
        if not ws.save:
            raise MixerClipError("backward needs a forward run with save=True")
        B, D, P, L, E, eng = ws.B, self.D, self.P, self.L, self.E, prec.engine
        st, G = self.store, self.store.grad_view
        s = ws.alloc_bwd(P, D, E, prec.act, st.device)
        dcur, dcur_a, dtmp = s["dcur"], s["dcur_a"], s["dtmp"]
        sep = prec.act != torch.float32
        if self.kind == "image":
            gam, proj, ln = self.p["visual.ln_post.weight"], "visual.proj", "visual.ln_post"
        else:
            gam, proj, ln = self.p["ln_final.weight"], "text_projection", "ln_final"
        # features: u = f/|f|                                                              model.py:433-434
        ops.l2norm_bwd(du_feat.contiguous(), ws.u_feat, ws.inv_norm, s["dfeat"], None, B, E)
        # f = a @ proj: dproj[D,E] += a^T dfeat ; da = dfeat proj^T   (fp32 FFMA engine, see forward)
        gp, ldgp = st.grad2d(proj)
        ops.gemm("simt", D, E, B, 1, ws.top_a, MAJOR_MN, D, 0, s["dfeat"], MAJOR_MN, E, 0, gp, ldgp, 0, accumulate=True)
        wp, ldp = st.weight_operand(proj, torch.float32)
        ops.gemm("simt", B, D, E, 1, s["dfeat"], MAJOR_K, E, 0, wp, MAJOR_K, ldp, 0, dtmp, D, 0)
        # only the cls / EOT row of the tower output carries gradient (Appendix A)
        dcur.zero_()
        if sep:
            dcur_a.zero_()
        gb4_last = G(f"{self.blk}.{L - 1}.channel_mix_seq.lin4.bias")
        ops.ln_bwd(dtmp, ws.x[L], D, ws.top_stats[0], ws.top_stats[1], gam, dcur, D, G(ln + ".weight"),
                   G(ln + ".bias"), B, D, row_index=ws.rows, dx_act=dcur_a if sep else None, colsum_out=gb4_last)
        if after_block:
            after_block("top")
        yield "top"
        for i in range(L - 1, -1, -1):
            self._block_bwd(i, ws, s, prec)
            if after_block:
                after_block(i)
            yield i
        if self.kind == "image":
            g = self.cfg["grid"]
            Kc = 3 * self.cfg["vision_patch_size"] ** 2
            pst = ws.extra["pre_stats"]
            dpre_a = s["dz1"].view(-1)[:B * P * D].view(B, P, D)          # reuse scratch for the operand copy
            ops.ln_bwd(dcur, ws.extra["xpre"], D, pst[0], pst[1], self.p["visual.ln_pre.weight"], None, D,
                       G("visual.ln_pre.weight"), G("visual.ln_pre.bias"), B * P, D,
                       cls=self.p["visual.class_embedding"], cls_period=P, dx_act=dpre_a,
                       dcls=G("visual.class_embedding"))
            gw, ldgw = st.grad2d("visual.conv1.weight")
            # dWconv[D, 3pp] += sum_b dpre[b, 1:, :]^T @ patches[b]   (reduction over batch x g*g)
            ops.gemm(eng, D, Kc, g * g, B, dpre_a.view(-1)[D:], MAJOR_MN, D, P * D, ws.extra["patches"], MAJOR_MN, Kc,
                     g * g * Kc, gw, ldgw, 0, k_spans_batch=True, accumulate=True, split_k=0)
        else:
            ops.embed_bwd(ws.extra["text"], dcur, G("token_embedding.weight"), B, self.cfg["context_length"], D,
                          self.cfg["vocab_size"])
        if after_block:
            after_block("bottom")
        yield "bottom"

    def _block_bwd(self, i, ws: TowerWS, s, prec: Precision):
        B, D, P, eng = ws.B, self.D, self.P, prec.engine
        st, G = self.store, self.store.grad_view
        sep = prec.act != torch.float32
        dcur, dcur_a, dtmp, dz2, dz1 = s["dcur"], s["dcur_a"], s["dtmp"], s["dz2"], s["dz1"]
        x, y, u, v = ws.x[i], ws.y[i], ws.u[i], ws.v[i]
        z2, h2, stt = ws.z2[i], ws.h2[i], ws.stats[i]
        h1 = ws.h1[i] if ws.h1 is not None else None
        pre = f"{self.blk}.{i}."
        T = B * P
        # ---- channel mix:  O = Y + g(V W3^T + b3) W4^T + b4 ----
        w4, ld4 = self.wop(pre + "channel_mix_seq.lin4.weight", prec)                 # [D, 4D]
        ops.gemm(eng, T, 4 * D, D, 1, dcur_a, MAJOR_K, D, 0, w4, MAJOR_MN, ld4, 0, dz2, 4 * D, 0, act=ACT_GELU_BWD,
                 zin=z2, ldzin=4 * D)                                                  # dZ2 = (dO W4) * g'(Z2)
        g4, ldg4 = st.grad2d(pre + "channel_mix_seq.lin4.weight")
        ops.gemm(eng, D, 4 * D, T, 1, dcur_a, MAJOR_MN, D, 0, h2, MAJOR_MN, 4 * D, 0, g4, ldg4, 0, accumulate=True,
                 split_k=0)                                                            # dW4 += dO^T H2
        g3, ldg3 = st.grad2d(pre + "channel_mix_seq.lin3.weight")
        ops.gemm(eng, 4 * D, D, T, 1, dz2, MAJOR_MN, 4 * D, 0, v, MAJOR_MN, D, 0, g3, ldg3, 0, accumulate=True,
                 split_k=0)                                                            # dW3 += dZ2^T V
        ops.colsum(dz2, T, 4 * D, 4 * D, G(pre + "channel_mix_seq.lin3.bias"))         # db3
        w3, ld3 = self.wop(pre + "channel_mix_seq.lin3.weight", prec)                 # [4D, D]
        ops.gemm(eng, T, D, 4 * D, 1, dz2, MAJOR_K, 4 * D, 0, w3, MAJOR_MN, ld3, 0, dtmp, D, 0)   # dV = dZ2 W3
        # dY = dO + LN2bwd(dV); db2[p] = sum_{b,d} dY
        ops.ln_bwd(dtmp, y, D, stt[2], stt[3], self.bp(i, "layerNorm2.weight"), dcur, D, G(pre + "layerNorm2.weight"),
                   G(pre + "layerNorm2.bias"), T, D, dres=dcur, dx_act=dcur_a if sep else None,
                   rowsum_out=G(pre + "token_mix_seq.lin2.bias"), rowsum_period=P)
        # ---- token mix:  Y = X + W2 g(W1 U + b1) + b2   (per sample, U = LN1(X)) ----
        w2, ld2 = self.wop(pre + "token_mix_seq.lin2.weight", prec)                   # [P, 4P]
        w1, ld1 = self.wop(pre + "token_mix_seq.lin1.weight", prec)                   # [4P, P]
        if ws.fused_tm:
            # weight gradients and the data gradient each in one kernel; both recompute Z1 / H1 / dZ1 per tile on chip
            g1, ldg1 = st.grad2d(pre + "token_mix_seq.lin1.weight")
            g2, ldg2 = st.grad2d(pre + "token_mix_seq.lin2.weight")
            b1 = self.bp(i, "token_mix_seq.lin1.bias")
            ops.token_mix_wgrad(B, P, D, u, dcur_a, w1, ld1, b1, w2, ld2, g1, ldg1, g2, ldg2,
                                G(pre + "token_mix_seq.lin1.bias"), w1t=self.w1t[i], ld1t=self.ld1t)
            ops.token_mix_dgrad(B, P, D, u, dcur_a, dtmp, w1, ld1, b1, w2, ld2, w1t=self.w1t[i], ld1t=self.ld1t)
        elif prec.recompute_z1:
            # dZ1 = (W2^T dY) * g'(W1 U + b1): the pre-activation is recomputed by a second operand pair of the same
            # GEMM tile (K = P is tiny), nothing was saved for it in the forward pass
            ops.gemm(eng, 4 * P, D, P, B, w2, MAJOR_MN, ld2, 0, dcur_a, MAJOR_MN, D, P * D, dz1, D, 4 * P * D,
                     act=ACT_GELU_BWD, rowsum_out=G(pre + "token_mix_seq.lin1.bias"),
                     A2=w1, a2_major=MAJOR_K, lda2=ld1, a2_bs=0, B2=u, b2_major=MAJOR_MN, ldb2=D, b2_bs=P * D,
                     bias2=self.bp(i, "token_mix_seq.lin1.bias"))
        else:
            ops.gemm(eng, 4 * P, D, P, B, w2, MAJOR_MN, ld2, 0, dcur_a, MAJOR_MN, D, P * D, dz1, D, 4 * P * D,
                     act=ACT_GELU_BWD, zin=ws.z1[i], ldzin=D, zin_bs=4 * P * D,
                     rowsum_out=G(pre + "token_mix_seq.lin1.bias"))                    # dZ1 = (W2^T dY) * g'(Z1); db1 fused
        if not ws.fused_tm:
            g2, ldg2 = st.grad2d(pre + "token_mix_seq.lin2.weight")
            ops.gemm(eng, P, 4 * P, D, B, dcur_a, MAJOR_K, D, P * D, h1, MAJOR_K, D, 4 * P * D, g2, ldg2, 0,
                     k_spans_batch=True, accumulate=True, split_k=0)                       # dW2 += sum_b dY H1^T
            g1, ldg1 = st.grad2d(pre + "token_mix_seq.lin1.weight")
            ops.gemm(eng, 4 * P, P, D, B, dz1, MAJOR_K, D, 4 * P * D, u, MAJOR_K, D, P * D, g1, ldg1, 0,
                     k_spans_batch=True, accumulate=True, split_k=0)                       # dW1 += sum_b dZ1 U^T
            ops.gemm(eng, P, D, 4 * P, B, w1, MAJOR_MN, ld1, 0, dz1, MAJOR_MN, D, 4 * P * D, dtmp, D, P * D)  # dU = W1^T dZ1
        # dX = dY + LN1bwd(dU); column sums of dX are db4 of the previous block
        prev_b4 = G(f"{self.blk}.{i - 1}.channel_mix_seq.lin4.bias") if i > 0 else None
        ops.ln_bwd(dtmp, x, D, stt[0], stt[1], self.bp(i, "layerNorm1.weight"), dcur, D, G(pre + "layerNorm1.weight"),
                   G(pre + "layerNorm1.bias"), T, D, dres=dcur, dx_act=dcur_a if sep else None, colsum_out=prev_b4)
