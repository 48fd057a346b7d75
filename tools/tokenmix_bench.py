#!/usr/bin/env python
"""Times the fused token-mixing kernels (mc_token_mix_fwd / _dgrad / _wgrad) on the two production towers
(CUDA events; the iterations rotate over buffer sets larger than L2) and prints achieved HBM GB/s against the
ALGORITHMIC bytes of DESIGN.md: fwd 10*P*D B per sample (u bf16 in, x fp32 in, y fp32 out), dgrad 8*P*D
(u, dy bf16 in, du fp32 out), wgrad 4*P*D (u, dy bf16 in; the outputs are two small weight matrices)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from clip_mixer_b200 import ops  # noqa: E402


def run(B, P, D, iters, nbuf, which):
    dev = torch.device("cuda:0")
    H = 4 * P
    ld1, ld2 = (P + 7) // 8 * 8, (H + 7) // 8 * 8
    u = torch.randn(nbuf, B, P, D, device=dev).to(torch.bfloat16)
    dy = torch.randn(nbuf, B, P, D, device=dev).to(torch.bfloat16)
    x = torch.randn(nbuf, B, P, D, device=dev)
    y = torch.empty(nbuf, B, P, D, device=dev)
    w1 = (torch.randn(H, ld1, device=dev) / P ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(P, ld2, device=dev) / H ** 0.5).to(torch.bfloat16)
    b1, b2 = torch.randn(H, device=dev), torch.randn(P, device=dev)
    w1t, ld1t = ops.w1_transposed(w1, P, ld1)          # refreshed once per step by the engine, not per launch
    gw1, gw2, gb1 = torch.zeros(H, ld1, device=dev), torch.zeros(P, ld2, device=dev), torch.zeros(H, device=dev)
    out = {}

    def fwd(i):
        ops.token_mix_fwd(B, P, D, u[i], x[i], y[i], w1, ld1, b1, w2, ld2, b2, w1t=w1t, ld1t=ld1t)

    # LayerNorm in the prologue: no bf16 operand in, the kernel writes it (u_out) plus mean / rstd
    sums = torch.stack([x.sum(-1), (x * x).sum(-1)], dim=-1).reshape(nbuf, B * P, 2).contiguous()
    gam, bet = torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev)
    mean, rstd = torch.empty(B * P, device=dev), torch.empty(B * P, device=dev)

    def fwd_ln(i):
        ops.token_mix_fwd(B, P, D, None, x[i], y[i], w1, ld1, b1, w2, ld2, b2, w1t=w1t, ld1t=ld1t,
                          ln=dict(sums=sums[i], gamma=gam, beta=bet, u_out=u[i], mean=mean, rstd=rstd))

    def ln_then_fwd(i):
        ops.ln_fwd(x[i], D, gam, bet, u[i], D, mean, rstd, B * P, D)
        fwd(i)

    def dgrad(i):
        ops.token_mix_dgrad(B, P, D, u[i], dy[i], y[i], w1, ld1, b1, w2, ld2, w1t=w1t, ld1t=ld1t)

    def wgrad(i):
        ops.token_mix_wgrad(B, P, D, u[i], dy[i], w1, ld1, b1, w2, ld2, gw1, ld1, gw2, ld2, gb1, w1t=w1t, ld1t=ld1t)

    # fwd_ln / ln+fwd: x fp32 in, u bf16 out, y fp32 out (+ the unfused pair re-reads u and x)
    for name, fn, bytes_per in (("fwd", fwd, 10), ("fwd_ln", fwd_ln, 10), ("ln+fwd", ln_then_fwd, 10), ("dgrad", dgrad, 8),
                                ("wgrad", wgrad, 4)):
        if which and name not in which:
            continue
        for i in range(3):
            fn(i % nbuf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # park the GPU on a spin kernel while the host enqueues the launches (each costs tens of microseconds of
        # Python / ctypes / tensor-map encoding): the events then bracket back-to-back kernel execution only
        torch.cuda._sleep(int(2e7))
        e0.record()
        for i in range(iters):
            fn(i % nbuf)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        alg = bytes_per * P * D * B
        out[name] = dict(us=round(us, 2), alg_MB=round(alg / 1e6, 1), GBps=round(alg / us / 1e3, 1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--nbuf", type=int, default=4)
    ap.add_argument("--only", default="")
    ap.add_argument("--tower", default="")
    a = ap.parse_args()
    which = [w for w in a.only.split(",") if w]
    res = {}
    for tag, P, D in (("image_P50_D768", 50, 768), ("text_P77_D512", 77, 512)):
        if a.tower and not tag.startswith(a.tower):
            continue
        res[tag] = run(a.batch, P, D, a.iters, a.nbuf, which)
    print(json.dumps(res))
    for tag, r in res.items():
        for k, v in r.items():
            print(f"{tag:16s} {k:6s} {v['us']:8.2f} us   {v['alg_MB']:7.1f} MB algorithmic   {v['GBps']:8.1f} GB/s")


if __name__ == "__main__":
    main()
