"""Thin tensor-level wrappers over the C ABI (include/mixerclip.h).

Every function takes CUDA tensors, passes raw device pointers + the current CUDA stream to
libmixerclip.so and returns nothing: outputs are written into caller-provided tensors (the
library never allocates).  A non-CUDA tensor raises; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_GELU_BWD, ACT_NONE, BF16, BIAS_M, BIAS_N, BIAS_NONE, F32, MAJOR_K, MAJOR_MN,
                   GemmParams, MixerClipError, TokenMixParams, check)

__all__ = ["gemm", "set_sm_limit", "token_mix_supported", "token_mix_fwd", "token_mix_dgrad", "token_mix_wgrad", "transpose_bf16", "w1_transposed", "ln_fwd", "ln_bwd", "colsum", "rowsum", "cast_pad", "im2col", "embed_fwd", "embed_bwd", "eot_rows",
           "l2norm_fwd", "l2norm_bwd", "head_fwd_bwd", "head_workspace_bytes", "sumsq", "sched_step", "adamw", "device_info",
           "F32", "BF16", "MAJOR_K", "MAJOR_MN", "BIAS_NONE", "BIAS_N", "BIAS_M", "ACT_NONE", "ACT_GELU",
           "ACT_GELU_BWD", "launch_count", "reset_launch_count", "enable_gemm_timing", "collect_gemm_timing"]

_DT = {torch.float32: F32, torch.bfloat16: BF16}
_launches = 0


def launch_count() -> int:
    """Number of libmixerclip kernel-launching calls made so far (bench.py's gpu_launches)."""
    return _launches


def reset_launch_count():
    global _launches
    _launches = 0


def _count(n=1):
    global _launches
    _launches += n


_gemm_timing = None


def enable_gemm_timing(on: bool):
    """Instrumentation for bench.py: bracket every GEMM launch with CUDA events on the launching stream."""
    global _gemm_timing
    _gemm_timing = [] if on else None


def collect_gemm_timing():
    """[{engine, M, N, K, batch, flops, ms}] for the launches since enable_gemm_timing(True); synchronises."""
    torch.cuda.synchronize()
    out = []
    for rec in (_gemm_timing or []):
        out.append(dict(engine=rec[0], M=rec[1], N=rec[2], K=rec[3], batch=rec[4], flops=2.0 * rec[1] * rec[2] * rec[3] * rec[4],
                        ms=rec[5].elapsed_time(rec[6]), tag=rec[7]))
    if _gemm_timing is not None:
        _gemm_timing.clear()
    return out


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise MixerClipError("libmixerclip ops need CUDA tensors (there is no CPU fallback)")
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise MixerClipError(f"unsupported dtype {t.dtype}") from None


def set_sm_limit(sms: int):
    """Persistent kernels launched after this call size their grids for ``sms`` SMs (0 = all): mc_set_sm_limit."""
    check(_lib.load().mc_set_sm_limit(int(sms)), "mc_set_sm_limit")


def device_info():
    lib = _lib.load()
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    check(lib.mc_device_info(C.byref(a), C.byref(b), C.byref(c)), "mc_device_info")
    return a.value, b.value, c.value


def gemm(engine: str, M: int, N: int, K: int, batch: int,
         A: torch.Tensor, a_major: int, lda: int, a_bs: int,
         B: torch.Tensor, b_major: int, ldb: int, b_bs: int,
         Cout: torch.Tensor, ldc: int, c_bs: int = 0, *,
         k_spans_batch: bool = False, accumulate: bool = False, split_k: int = 1, row_remap: int = 0,
         bias: Optional[torch.Tensor] = None, bias_mode: int = BIAS_NONE,
         zout: Optional[torch.Tensor] = None, ldz: int = 0, z_bs: int = 0,
         zin: Optional[torch.Tensor] = None, ldzin: int = 0, zin_bs: int = 0,
         act: int = ACT_NONE, R: Optional[torch.Tensor] = None, ldr: int = 0, r_bs: int = 0,
         rowsum_out: Optional[torch.Tensor] = None, c_transposed: bool = False,
         A2: Optional[torch.Tensor] = None, a2_major: int = 0, lda2: int = 0, a2_bs: int = 0,
         B2: Optional[torch.Tensor] = None, b2_major: int = 0, ldb2: int = 0, b2_bs: int = 0,
         bias2: Optional[torch.Tensor] = None, rowstat_out: Optional[torch.Tensor] = None):
    """acc[m,n] = sum_k A[b][m,k] B[b][n,k] with the fused epilogue of mc_gemm_params.

    engine "tc"   -> mc_gemm_bf16_tc  (A, B, zout, zin bf16)
    engine "simt" -> mc_gemm_f32_simt (everything fp32)
    """
    lib = _lib.load()
    want = torch.bfloat16 if engine == "tc" else torch.float32
    zwant = torch.float16 if engine == "tc" else torch.float32
    for name, t, w in (("A", A, want), ("B", B, want), ("zout", zout, zwant), ("zin", zin, zwant)):
        if t is not None and t.dtype != w:
            raise MixerClipError(f"gemm[{engine}]: operand {name} must be {w}, got {t.dtype}")
    for name, t in (("bias", bias), ("R", R)):
        if t is not None and t.dtype != torch.float32:
            raise MixerClipError(f"gemm: {name} must be fp32")
    p = GemmParams()
    p.M, p.N, p.K, p.batch = M, N, K, batch
    p.A, p.a_major, p.lda, p.a_batch_stride = _ptr(A), a_major, lda, a_bs
    p.B, p.b_major, p.ldb, p.b_batch_stride = _ptr(B), b_major, ldb, b_bs
    p.k_spans_batch = 1 if k_spans_batch else 0
    p.C, p.c_dtype, p.ldc, p.c_batch_stride = _ptr(Cout), dtype_code(Cout), ldc, c_bs
    p.accumulate, p.split_k, p.row_remap = (1 if accumulate else 0), split_k, row_remap
    p.bias, p.bias_mode = _ptr(bias), bias_mode
    p.zout, p.ldz, p.z_batch_stride = _ptr(zout), ldz, z_bs
    p.zin, p.ldzin, p.zin_batch_stride = _ptr(zin), ldzin, zin_bs
    p.act = act
    p.R, p.ldr, p.r_batch_stride = _ptr(R), ldr, r_bs
    p.rowsum_out = _ptr(rowsum_out)
    p.c_transposed = 1 if c_transposed else 0
    for name, t in (("A2", A2), ("B2", B2)):
        if t is not None and t.dtype != want:
            raise MixerClipError(f"gemm[{engine}]: operand {name} must be {want}, got {t.dtype}")
    p.A2, p.a2_major, p.lda2, p.a2_batch_stride = _ptr(A2), a2_major, lda2, a2_bs
    p.B2, p.b2_major, p.ldb2, p.b2_batch_stride = _ptr(B2), b2_major, ldb2, b2_bs
    p.bias2 = _ptr(bias2)
    if rowstat_out is not None and rowstat_out.dtype != torch.float32:
        raise MixerClipError("gemm: rowstat_out must be fp32")
    p.rowstat_out = _ptr(rowstat_out)
    fn = lib.mc_gemm_bf16_tc if engine == "tc" else lib.mc_gemm_f32_simt
    if _gemm_timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(fn(C.byref(p), _stream()), f"mc_gemm[{engine}] M={M} N={N} K={K} batch={batch}")
    if _gemm_timing is not None:
        e1.record()
        tag = f"a{a_major}b{b_major}" + ("kb" if k_spans_batch else "") + (f"act{act}" if act else "")
        _gemm_timing.append((engine, M, N, K, batch, e0, e1, tag))
    _count()


def token_mix_supported(P: int, D: int) -> bool:
    """True when the fused token-mixing kernels (mc_token_mix_*) cover the shape (P <= 80 tokens, D % 128 == 0)."""
    return bool(_lib.load().mc_token_mix_supported(P, D))


def transpose_bf16(src, rows, cols, ld_src, src_bs, dst, ld_dst, dst_bs, batch=1):
    """dst[b][c][r] = src[b][r][c] for `batch` bf16 matrices (W1^T copies of the token-mixing lin1 weights)."""
    if src.dtype != torch.bfloat16 or dst.dtype != torch.bfloat16:
        raise MixerClipError("transpose_bf16: bf16 tensors only")
    check(_lib.load().mc_transpose_bf16(_ptr(src), rows, cols, ld_src, src_bs, _ptr(dst), ld_dst, dst_bs, batch, _stream()),
          "mc_transpose_bf16")
    _count()


def w1_transposed(w1, P, ld1):
    """Stand-alone W1^T [P, ld1t] of one lin1 weight operand [4P, ld1] (tests / tools; the engine keeps per-tower copies)."""
    H = 4 * P
    ld1t = (H + 7) // 8 * 8
    out = torch.empty(P, ld1t, device=w1.device, dtype=torch.bfloat16)
    transpose_bf16(w1, H, P, ld1, 0, out, ld1t, 0, 1)
    return out, ld1t


def _tm_params(B, P, D, u, w1, ld1, b1, w2, ld2, w1t=None, ld1t=0):
    for name, t in (("u", u), ("w1", w1), ("w2", w2)):
        if t is not None and t.dtype != torch.bfloat16:
            raise MixerClipError(f"token_mix: operand {name} must be bf16, got {t.dtype}")
    if w1t is None:
        w1t, ld1t = w1_transposed(w1, P, ld1)
    p = TokenMixParams()
    p.B, p.P, p.D = B, P, D
    p.u, p.w1, p.ld1, p.b1, p.w2, p.ld2 = _ptr(u), _ptr(w1), ld1, _ptr(b1), _ptr(w2), ld2
    p.w1t, p.ld1t = _ptr(w1t), ld1t
    p._keep = w1t            # keep the transposed copy alive until the (asynchronous) launch has been enqueued
    return p


def _tm_call(fn, p, what, B, P, D):
    if _gemm_timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(fn(C.byref(p), _stream()), what)
    if _gemm_timing is not None:
        e1.record()
        # FLOPs executed on the tensor cores: 2 GEMMs (fwd), 3 (dgrad), 4 (wgrad) of 2*4P*P*D per sample
        n = {"token_mix_fwd": 2, "token_mix_dgrad": 3, "token_mix_wgrad": 4}[what]
        _gemm_timing.append((what, 4 * P, D, P, B * n, e0, e1, what))
    _count()


def token_mix_fwd(B, P, D, u, x, y, w1, ld1, b1, w2, ld2, b2, w1t=None, ld1t=0, ln=None):
    """y = x + W2 g(W1 u + b1) + b2 per sample (model.py:216,220-222), one fused kernel.

    ln = dict(sums [B*P,2] fp32, gamma [D], beta [D], u_out [B,P,D] bf16, mean [B*P], rstd [B*P]): LayerNorm in the prologue -
    the kernel computes u = LN(x) itself from the row sums of the producing GEMM (gemm(..., rowstat_out=sums)); `u` is ignored."""
    p = _tm_params(B, P, D, None if ln is not None else u, w1, ld1, b1, w2, ld2, w1t, ld1t)
    p.b2, p.x, p.y = _ptr(b2), _ptr(x), _ptr(y)
    if ln is not None:
        if ln["u_out"].dtype != torch.bfloat16 or any(ln[k].dtype != torch.float32 for k in ("sums", "gamma", "beta", "mean", "rstd")):
            raise MixerClipError("token_mix_fwd: ln = {sums, gamma, beta, mean, rstd: fp32; u_out: bf16}")
        p.ln_sums, p.ln_gamma, p.ln_beta = _ptr(ln["sums"]), _ptr(ln["gamma"]), _ptr(ln["beta"])
        p.u_out, p.ln_mean, p.ln_rstd = _ptr(ln["u_out"]), _ptr(ln["mean"]), _ptr(ln["rstd"])
    _tm_call(_lib.load().mc_token_mix_fwd, p, "token_mix_fwd", B, P, D)


def token_mix_dgrad(B, P, D, u, dy, du, w1, ld1, b1, w2, ld2, w1t=None, ld1t=0):
    """du = W1^T ((W2^T dy) * g'(W1 u + b1)) per sample, one fused kernel (dy bf16, du fp32)."""
    if dy.dtype != torch.bfloat16 or du.dtype != torch.float32:
        raise MixerClipError("token_mix_dgrad: dy must be bf16 and du fp32")
    p = _tm_params(B, P, D, u, w1, ld1, b1, w2, ld2, w1t, ld1t)
    p.dy, p.y = _ptr(dy), _ptr(du)
    _tm_call(_lib.load().mc_token_mix_dgrad, p, "token_mix_dgrad", B, P, D)


def token_mix_wgrad(B, P, D, u, dy, w1, ld1, b1, w2, ld2, gw1, ldg1, gw2, ldg2, gb1, w1t=None, ld1t=0):
    """gw1 += sum_b dZ1 u^T, gw2 += sum_b dy H1^T, gb1 += rowsum(dZ1); H1 and dZ1 are recomputed on chip."""
    if dy.dtype != torch.bfloat16:
        raise MixerClipError("token_mix_wgrad: dy must be bf16")
    for t in (gw1, gw2, gb1):
        if t.dtype != torch.float32:
            raise MixerClipError("token_mix_wgrad: gradients must be fp32")
    p = _tm_params(B, P, D, u, w1, ld1, b1, w2, ld2, w1t, ld1t)
    p.dy, p.gw1, p.ldg1, p.gw2, p.ldg2, p.gb1 = _ptr(dy), _ptr(gw1), ldg1, _ptr(gw2), ldg2, _ptr(gb1)
    _tm_call(_lib.load().mc_token_mix_wgrad, p, "token_mix_wgrad", B, P, D)


def ln_fwd(x, x_row_stride, gamma, beta, y, y_row_stride, mean, rstd, rows, D, *, row_index=None, cls=None,
           cls_period=0):
    check(_lib.load().mc_ln_fwd(_ptr(x), x_row_stride, _ptr(row_index), _ptr(cls), cls_period, _ptr(gamma), _ptr(beta),
                                _ptr(y), dtype_code(y), y_row_stride, _ptr(mean), _ptr(rstd), rows, D, _stream()),
          "mc_ln_fwd")
    _count()


def ln_bwd(dy, x, x_row_stride, mean, rstd, gamma, dx, dx_row_stride, dgamma, dbeta, rows, D, *, row_index=None,
           cls=None, cls_period=0, dres=None, dx_act=None, colsum_out=None, rowsum_out=None, rowsum_period=0,
           dcls=None):
    act_code = dtype_code(dx_act) if dx_act is not None else F32
    check(_lib.load().mc_ln_bwd(_ptr(dy), _ptr(x), x_row_stride, _ptr(row_index), _ptr(cls), cls_period, _ptr(mean),
                                _ptr(rstd), _ptr(gamma), _ptr(dres), _ptr(dx), dx_row_stride, _ptr(dx_act), act_code,
                                _ptr(dgamma), _ptr(dbeta), _ptr(colsum_out), _ptr(rowsum_out), rowsum_period,
                                _ptr(dcls), rows, D, _stream()), "mc_ln_bwd")
    _count()


def colsum(x, rows, cols, ld, out):
    check(_lib.load().mc_colsum(_ptr(x), dtype_code(x), rows, cols, ld, _ptr(out), _stream()), "mc_colsum")
    _count()


def rowsum(x, rows, cols, ld, period, out):
    check(_lib.load().mc_rowsum(_ptr(x), dtype_code(x), rows, cols, ld, period, _ptr(out), _stream()), "mc_rowsum")
    _count()


def cast_pad(src, rows, cols, src_ld, dst, dst_ld):
    check(_lib.load().mc_cast_pad(_ptr(src), rows, cols, src_ld, _ptr(dst), dtype_code(dst), dst_ld, _stream()),
          "mc_cast_pad")
    _count()


def im2col(image, B, R, patch, out):
    if image.dtype not in (torch.float32, torch.uint8):
        raise MixerClipError(f"im2col: image must be fp32 or uint8, got {image.dtype}")
    check(_lib.load().mc_im2col(_ptr(image), 1 if image.dtype == torch.uint8 else 0, B, R, patch, _ptr(out),
                                dtype_code(out), _stream()), "mc_im2col")
    _count()


def embed_fwd(text, table, x, B, Cn, W, vocab):
    check(_lib.load().mc_embed_fwd(_ptr(text), _ptr(table), _ptr(x), B, Cn, W, vocab, _stream()), "mc_embed_fwd")
    _count()


def embed_bwd(text, dx, dtable, B, Cn, W, vocab):
    check(_lib.load().mc_embed_bwd(_ptr(text), _ptr(dx), _ptr(dtable), B, Cn, W, vocab, _stream()), "mc_embed_bwd")
    _count()


def eot_rows(text, out, B, Cn):
    check(_lib.load().mc_eot_rows(_ptr(text), _ptr(out), B, Cn, _stream()), "mc_eot_rows")
    _count()


def l2norm_fwd(f, u, inv_norm, rows, E):
    check(_lib.load().mc_l2norm_fwd(_ptr(f), _ptr(u), _ptr(inv_norm), rows, E, _stream()), "mc_l2norm_fwd")
    _count()


def l2norm_bwd(du, u, inv_norm, df, df_act, rows, E):
    act_code = dtype_code(df_act) if df_act is not None else F32
    check(_lib.load().mc_l2norm_bwd(_ptr(du), _ptr(u), _ptr(inv_norm), _ptr(df), _ptr(df_act), act_code, rows, E,
                                    _stream()), "mc_l2norm_bwd")
    _count()


def head_workspace_bytes(n, N, E) -> int:
    return int(_lib.load().mc_head_workspace_bytes(n, N, E))


def head_fwd_bwd(ui, ut, ui_all, ut_all, log_scale, n, N, E, rank, grad_scale, loss, dui, dut, dlog_scale, workspace):
    check(_lib.load().mc_head_fwd_bwd(_ptr(ui), _ptr(ut), _ptr(ui_all), _ptr(ut_all), _ptr(log_scale), n, N, E, rank,
                                      float(grad_scale), _ptr(loss), _ptr(dui), _ptr(dut), _ptr(dlog_scale),
                                      _ptr(workspace), workspace.numel() * workspace.element_size(), _stream()),
          "mc_head_fwd_bwd")
    _count(2)


def sumsq(g, n, out):
    check(_lib.load().mc_sumsq(_ptr(g), n, _ptr(out), _stream()), "mc_sumsq")
    _count()


def sched_step(state, hyper, first_cycle_steps, max_lr, min_lr, warmup_steps, beta1, beta2, fixed_lr=-1.0):
    """Device-side scheduler + Adam step count: hyper = {lr, 1-b1^t, 1-b2^t}, state (int64 {t, s}) advanced."""
    if state.dtype != torch.int64 or hyper.dtype != torch.float32:
        raise MixerClipError("sched_step: state must be int64[2] and hyper fp32[3]")
    check(_lib.load().mc_sched_step(_ptr(state), _ptr(hyper), int(first_cycle_steps), float(max_lr), float(min_lr),
                                    int(warmup_steps), float(beta1), float(beta2), float(fixed_lr), _stream()),
          "mc_sched_step")
    _count()


def adamw(p, g, m, v, p_bf16, decay_flags, n, sumsq_buf, hyper, grad_mul, max_norm, beta1, beta2, eps, weight_decay):
    check(_lib.load().mc_adamw(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p_bf16), _ptr(decay_flags), n, _ptr(sumsq_buf),
                               _ptr(hyper), float(grad_mul), float(max_norm), float(beta1), float(beta2), float(eps),
                               float(weight_decay), _stream()), "mc_adamw")
    _count()
