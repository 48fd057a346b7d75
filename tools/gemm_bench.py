#!/usr/bin/env python
"""Times individual GEMM shapes of the Mixer-CLIP step through the C ABI (CUDA events, L2 flushed between
iterations by rotating over buffers larger than L2) - also the target of `ncu --set full` captures."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from clip_mixer_b200 import ops  # noqa: E402

B, P, D = 256, 50, 768
SHAPES = {
    # name: (M, N, K, batch, a_major, b_major, kind)
    "lin3": (B * P, 4 * D, D, 1, 0, 0, "act_fwd"),
    "lin4": (B * P, D, 4 * D, 1, 0, 0, "resid"),
    "dz2": (B * P, 4 * D, D, 1, 0, 1, "act_bwd"),
    "dv": (B * P, D, 4 * D, 1, 0, 1, "plain"),
    "dw3": (4 * D, D, B * P, 1, 1, 1, "acc"),
    "dw4": (D, 4 * D, B * P, 1, 1, 1, "acc"),
    "tok_lin1": (4 * P, D, P, B, 0, 1, "tok_act_fwd"),
    "tok_lin2": (P, D, 4 * P, B, 0, 1, "tok_resid"),
    "tok_dz1": (4 * P, D, P, B, 1, 1, "tok_act_bwd"),
    "tok_du": (P, D, 4 * P, B, 1, 1, "tok_plain"),
    "txt_lin3": (B * 77, 2048, 512, 1, 0, 0, "act_fwd"),
    "txt_lin4": (B * 77, 512, 2048, 1, 0, 0, "resid"),
    "txt_dz2": (B * 77, 2048, 512, 1, 0, 1, "act_bwd"),
    "txt_dv": (B * 77, 512, 2048, 1, 0, 1, "plain"),
    "txt_dw3": (2048, 512, B * 77, 1, 1, 1, "acc"),
    "txt_dw4": (512, 2048, B * 77, 1, 1, 1, "acc"),
}


def make_operand(rows, K, batch, major, shared, dev, nbuf):
    ld = (K + 7) // 8 * 8 if major == 0 else (rows + 7) // 8 * 8
    outer = rows if major == 0 else K
    nb = 1 if shared else batch
    t = torch.randn(nbuf, nb, outer, ld, device=dev).to(torch.bfloat16)
    return t, ld, (0 if shared else outer * ld)


def run(name, iters, nbuf):
    M, N, K, batch, am, bm, kind = SHAPES[name]
    dev = torch.device("cuda:0")
    tok = kind.startswith("tok_")
    A, lda, a_bs = make_operand(M, K, batch, am, tok, dev, nbuf)
    Bm, ldb, b_bs = make_operand(N, K, batch, bm, False, dev, nbuf)
    k = kind.replace("tok_", "")
    cdt = torch.bfloat16 if k in ("act_fwd", "act_bwd") else torch.float32
    C = torch.zeros(nbuf, batch, M, N, device=dev, dtype=cdt)
    Z = torch.zeros(nbuf, batch, M, N, device=dev, dtype=torch.float16) if k in ("act_fwd", "act_bwd") else None
    R = torch.randn(nbuf, batch, M, N, device=dev) if k == "resid" else None
    bias = torch.randn(max(M, N), device=dev)
    kw = {}
    if k == "act_fwd":
        kw = dict(bias=bias, bias_mode=ops.BIAS_M if tok else ops.BIAS_N, act=ops.ACT_GELU, ldz=N, z_bs=M * N)
    elif k == "resid":
        kw = dict(bias=bias, bias_mode=ops.BIAS_M if tok else ops.BIAS_N, ldr=N, r_bs=M * N)
    elif k == "act_bwd":
        kw = dict(act=ops.ACT_GELU_BWD, ldzin=N, zin_bs=M * N)
    elif k == "acc":
        kw = dict(accumulate=True, split_k=0)

    def call(i):
        j = i % nbuf
        extra = dict(kw)
        if k == "act_fwd":
            extra["zout"] = Z[j]
        if k == "act_bwd":
            extra["zin"] = Z[j]
        if k == "resid":
            extra["R"] = R[j]
        ops.gemm("tc", M, N, K, batch, A[j], am, lda, a_bs, Bm[j], bm, ldb, b_bs, C[j], N, M * N, **extra)

    for i in range(3):
        call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        call(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * M * N * K * batch
    print(f"{name:10s} M={M} N={N} K={K} batch={batch} a{am}b{bm} {kind:12s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*", default=list(SHAPES))
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--nbuf", type=int, default=3)
    a = ap.parse_args()
    for n in a.names:
        run(n, a.iters, a.nbuf)
