#!/usr/bin/env python
"""Times the dense LayerNorm kernels of a Mixer block (mc_ln_fwd / mc_ln_bwd) on the two production towers with CUDA
events (buffers rotate over sets larger than L2) and prints achieved HBM GB/s against the algorithmic bytes of
DESIGN.md: forward 6*D B per row (fp32 in, bf16 out), backward 18*D B per row (dy, x, dres in; dx fp32 + bf16 out)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from clip_mixer_b200 import ops  # noqa: E402


def run(rows, D, P, iters, nbuf):
    dev = torch.device("cuda:0")
    x = torch.randn(nbuf, rows, D, device=dev)
    dy = torch.randn(nbuf, rows, D, device=dev)
    dres = torch.randn(nbuf, rows, D, device=dev)
    dx = torch.empty(nbuf, rows, D, device=dev)
    dxa = torch.empty(nbuf, rows, D, device=dev, dtype=torch.bfloat16)
    y = torch.empty(nbuf, rows, D, device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.zeros(rows, device=dev), torch.ones(rows, device=dev)
    gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    dg, db, cs = (torch.zeros(D, device=dev) for _ in range(3))
    rsum = torch.zeros(P, device=dev)

    def fwd(i):
        ops.ln_fwd(x[i], D, gamma, beta, y[i], D, mean, rstd, rows, D)

    def bwd(i):
        ops.ln_bwd(dy[i], x[i], D, mean, rstd, gamma, dx[i], D, dg, db, rows, D, dres=dres[i], dx_act=dxa[i], colsum_out=cs)

    def bwd_rs(i):
        ops.ln_bwd(dy[i], x[i], D, mean, rstd, gamma, dx[i], D, dg, db, rows, D, dres=dres[i], dx_act=dxa[i], rowsum_out=rsum,
                   rowsum_period=P)

    for name, fn, per in (("ln_fwd", fwd, 6), ("ln_bwd+colsum", bwd, 18), ("ln_bwd+rowsum", bwd_rs, 18)):
        for i in range(3):
            fn(i % nbuf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(2e7))      # the host enqueues while the GPU is parked: events bracket kernel execution only
        e0.record()
        for i in range(iters):
            fn(i % nbuf)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        alg = per * D * rows
        print(f"rows={rows} D={D} {name:14s} {us:8.2f} us  {alg / 1e6:7.1f} MB algorithmic  {alg / us / 1e3:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--nbuf", type=int, default=4)
    ap.add_argument("--tower", default="")
    a = ap.parse_args()
    if a.tower in ("", "image"):
        run(256 * 50, 768, 50, a.iters, a.nbuf)
    if a.tower in ("", "text"):
        run(256 * 77, 512, 77, a.iters, a.nbuf)
