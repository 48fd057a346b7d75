"""GPU parity of the row-wise kernels, the contrastive head and the optimizer through the C ABI.

Checkers: torch fp64 autograd of the same op (floating-point kernels), and the oracle's closed-form
head (oracle/mixer_clip_oracle.py).  Tolerance 1e-5 relative (fp32 math everywhere here) unless the
output is a bf16 operand copy (4e-3 = one bf16 rounding).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _ops():
    from clip_mixer_b200 import ops
    return ops


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def ln_ref(x, g, b):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, 1e-5)


@pytest.mark.parametrize("D", [768, 512, 64, 40, 1024])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_ln_fwd(D, out_dtype):
    ops = _ops()
    torch.manual_seed(0)
    rows = 203
    x = torch.randn(rows, D, device=DEV) * 2 + 0.5
    g = torch.randn(D, device=DEV)
    b = torch.randn(D, device=DEV)
    y = torch.empty(rows, D, device=DEV, dtype=out_dtype)
    mean = torch.empty(rows, device=DEV)
    rstd = torch.empty(rows, device=DEV)
    ops.ln_fwd(x, D, g, b, y, D, mean, rstd, rows, D)
    ref = ln_ref(x.double(), g.double(), b.double())
    assert rel(y, ref) < (1e-5 if out_dtype == torch.float32 else 4e-3)
    assert rel(mean, x.double().mean(1)) < 1e-5
    assert rel(rstd, 1 / torch.sqrt(x.double().var(1, unbiased=False) + 1e-5)) < 1e-5


def test_ln_fwd_cls_and_row_index():
    ops = _ops()
    torch.manual_seed(1)
    B, P, D = 5, 10, 64
    x = torch.randn(B, P, D, device=DEV)
    cls = torch.randn(D, device=DEV)
    g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    y = torch.empty(B, P, D, device=DEV)
    mean, rstd = torch.empty(B * P, device=DEV), torch.empty(B * P, device=DEV)
    ops.ln_fwd(x, D, g, b, y, D, mean, rstd, B * P, D, cls=cls, cls_period=P)
    xr = x.clone()
    xr[:, 0, :] = cls
    assert rel(y, ln_ref(xr.double(), g.double(), b.double())) < 1e-5
    # gather rows (EOT rows / class rows) through row_index
    idx = torch.tensor([3, 17, 49, 0, 22], device=DEV, dtype=torch.int32)
    y2 = torch.empty(5, D, device=DEV, dtype=torch.bfloat16)
    m2, r2 = torch.empty(5, device=DEV), torch.empty(5, device=DEV)
    ops.ln_fwd(x, D, g, b, y2, D, m2, r2, 5, D, row_index=idx)
    assert rel(y2, ln_ref(x.reshape(-1, D)[idx.long()].double(), g.double(), b.double())) < 4e-3


@pytest.mark.parametrize("D", [768, 512, 48])
def test_ln_bwd_full(D):
    ops = _ops()
    torch.manual_seed(2)
    B, P = 6, 11
    rows = B * P
    x = (torch.randn(rows, D, device=DEV) * 1.5).double().requires_grad_(True)
    g = torch.randn(D, device=DEV).double().requires_grad_(True)
    b = torch.randn(D, device=DEV).double().requires_grad_(True)
    dy = torch.randn(rows, D, device=DEV)
    dres = torch.randn(rows, D, device=DEV)
    y = ln_ref(x, g, b)
    (y * dy.double()).sum().backward()
    dx_ref = x.grad + dres.double()
    xf, gf = x.detach().float(), g.detach().float()
    mean = xf.double().mean(1).float()
    rstd = (1 / torch.sqrt(xf.double().var(1, unbiased=False) + 1e-5)).float()
    dx = torch.empty(rows, D, device=DEV)
    dxa = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
    dgamma, dbeta = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    cs, rs = torch.zeros(D, device=DEV), torch.zeros(P, device=DEV)
    ops.ln_bwd(dy, xf, D, mean, rstd, gf, dx, D, dgamma, dbeta, rows, D, dres=dres, dx_act=dxa, colsum_out=cs,
               rowsum_out=rs, rowsum_period=P)
    assert rel(dx, dx_ref) < 1e-5
    assert rel(dxa, dx_ref) < 4e-3
    assert rel(dgamma, g.grad) < 1e-5 and rel(dbeta, b.grad) < 1e-5
    assert rel(cs, dx_ref.sum(0)) < 1e-5
    assert rel(rs, dx_ref.reshape(B, P, D).sum((0, 2))) < 1e-5


def test_ln_bwd_cls_rows_and_scatter():
    ops = _ops()
    torch.manual_seed(3)
    B, P, D = 4, 5, 64
    rows = B * P
    xpre = torch.randn(B, P, D, device=DEV)
    cls = torch.randn(D, device=DEV)
    g = torch.randn(D, device=DEV)
    dy = torch.randn(rows, D, device=DEV)
    xr = xpre.clone()
    xr[:, 0, :] = cls
    x64 = xr.double().requires_grad_(True)
    y = ln_ref(x64, g.double(), torch.zeros(D, device=DEV).double())
    (y * dy.double().reshape(B, P, D)).sum().backward()
    mean = xr.double().mean(2).float().reshape(-1)
    rstd = (1 / torch.sqrt(xr.double().var(2, unbiased=False) + 1e-5)).float().reshape(-1)
    dxa = torch.full((B, P, D), 5.0, device=DEV)
    dgamma, dbeta, dcls = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.ln_bwd(dy, xpre, D, mean, rstd, g, None, D, dgamma, dbeta, rows, D, cls=cls, cls_period=P, dx_act=dxa,
               dcls=dcls)
    assert rel(dcls, x64.grad[:, 0, :].sum(0)) < 1e-5
    assert rel(dxa[:, 1:, :], x64.grad[:, 1:, :]) < 1e-5
    assert float(dxa[:, 0, :].abs().max()) == 0.0
    # scatter form (ln_post / ln_final backward): rows selected by row_index, dx written in place
    idx = torch.tensor([0, 7, 13, 19], device=DEV, dtype=torch.int32)
    dy2 = torch.randn(4, D, device=DEV)
    sel = xpre.reshape(-1, D)[idx.long()]
    s64 = sel.double().requires_grad_(True)
    (ln_ref(s64, g.double(), torch.zeros(D, device=DEV).double()) * dy2.double()).sum().backward()
    m2 = sel.double().mean(1).float()
    r2 = (1 / torch.sqrt(sel.double().var(1, unbiased=False) + 1e-5)).float()
    dfull = torch.zeros(B * P, D, device=DEV)
    dg2, db2, cs = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.ln_bwd(dy2, xpre, D, m2, r2, g, dfull, D, dg2, db2, 4, D, row_index=idx, colsum_out=cs)
    exp = torch.zeros(B * P, D, device=DEV, dtype=torch.float64)
    exp[idx.long()] = s64.grad
    assert rel(dfull, exp) < 1e-5
    assert rel(cs, s64.grad.sum(0)) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_colsum_rowsum(dtype):
    ops = _ops()
    torch.manual_seed(4)
    x = torch.randn(1000, 200, device=DEV).to(dtype)
    out = torch.ones(200, device=DEV)
    ops.colsum(x, 1000, 200, 200, out)
    assert rel(out, 1 + x.double().sum(0)) < 1e-5
    out2 = torch.zeros(8, device=DEV)
    ops.rowsum(x, 1000, 200, 200, 8, out2)
    assert rel(out2, x.double().sum(1).reshape(125, 8).sum(0)) < 1e-5


def test_cast_pad():
    ops = _ops()
    w = torch.randn(200, 50, device=DEV)
    dst = torch.full((200, 56), 3.0, device=DEV, dtype=torch.bfloat16)
    ops.cast_pad(w, 200, 50, 50, dst, 56)
    assert torch.equal(dst[:, :50], w.to(torch.bfloat16))
    assert float(dst[:, 50:].abs().max()) == 0.0


@pytest.mark.parametrize("patch,R", [(32, 224), (16, 64)])
def test_im2col(patch, R):
    ops = _ops()
    import sys
    from oracle import mixer_clip_oracle as O
    torch.manual_seed(5)
    B = 3
    img = torch.randn(B, 3, R, R, device=DEV)
    g = R // patch
    out = torch.empty(B * g * g, 3 * patch * patch, device=DEV)
    ops.im2col(img, B, R, patch, out)
    assert torch.equal(out, O.patchify(img, patch).reshape(B * g * g, -1))
    u8 = torch.randint(0, 256, (B, 3, R, R), device=DEV, dtype=torch.uint8)
    out16 = torch.empty(B * g * g, 3 * patch * patch, device=DEV, dtype=torch.bfloat16)
    ops.im2col(u8, B, R, patch, out16)
    mean = torch.tensor(O.IMAGE_MEAN, device=DEV).view(1, 3, 1, 1)
    std = torch.tensor(O.IMAGE_STD, device=DEV).view(1, 3, 1, 1)
    ref = O.patchify((u8.float() / 255 - mean) / std, patch).reshape(B * g * g, -1)
    assert rel(out16, ref) < 4e-3


def test_embedding_and_eot():
    ops = _ops()
    from oracle import mixer_clip_oracle as O
    cfg = O.CONFIGS["tiny"]
    torch.manual_seed(6)
    B, Cn, W, V = 9, 12, 48, 100
    _, text = O.synthetic_batch(cfg, B, seed=3)
    text = text.to(DEV)
    table = torch.randn(V, W, device=DEV)
    x = torch.empty(B, Cn, W, device=DEV)
    ops.embed_fwd(text, table, x, B, Cn, W, V)
    assert torch.equal(x, table[text])
    dx = torch.randn(B, Cn, W, device=DEV)
    dt = torch.zeros(V, W, device=DEV)
    ops.embed_bwd(text, dx, dt, B, Cn, W, V)
    ref = torch.zeros(V, W, device=DEV, dtype=torch.float64).index_add_(0, text.reshape(-1), dx.double().reshape(-1, W))
    assert rel(dt, ref) < 1e-5
    eot = torch.empty(B, device=DEV, dtype=torch.int32)
    ops.eot_rows(text, eot, B, Cn)
    assert torch.equal(eot.long(), torch.arange(B, device=DEV) * Cn + text.argmax(-1))


@pytest.mark.parametrize("E", [512, 32, 16])
def test_l2norm(E):
    ops = _ops()
    torch.manual_seed(7)
    n = 37
    f = torch.randn(n, E, device=DEV).double().requires_grad_(True)
    du = torch.randn(n, E, device=DEV)
    u_ref = f / f.norm(dim=1, keepdim=True)
    (u_ref * du.double()).sum().backward()
    u = torch.empty(n, E, device=DEV)
    inv = torch.empty(n, device=DEV)
    ops.l2norm_fwd(f.detach().float(), u, inv, n, E)
    assert rel(u, u_ref) < 1e-5
    df = torch.empty(n, E, device=DEV)
    dfa = torch.empty(n, E, device=DEV, dtype=torch.bfloat16)
    ops.l2norm_bwd(du, u, inv, df, dfa, n, E)
    assert rel(df, f.grad) < 1e-5 and rel(dfa, f.grad) < 4e-3


@pytest.mark.parametrize("n,N,E,rank", [(8, 8, 32, 0), (37, 37, 512, 0), (16, 64, 16, 2), (256, 256, 512, 0),
                                         (96, 384, 512, 3)])
def test_head_against_oracle_closed_form(n, N, E, rank):
    ops = _ops()
    from oracle import mixer_clip_oracle as O
    torch.manual_seed(8)
    ui_all = torch.nn.functional.normalize(torch.randn(N, E, dtype=torch.float64), dim=1)
    ut_all = torch.nn.functional.normalize(torch.randn(N, E, dtype=torch.float64) + 0.5 * ui_all, dim=1)
    ui, ut = ui_all[rank * n:(rank + 1) * n], ut_all[rank * n:(rank + 1) * n]
    t = torch.tensor(math.log(1 / 0.07), dtype=torch.float64)
    loss_ref, dui_ref, dut_ref, dt_ref = O.head_closed_form(ui, ut, t, ui_all, ut_all, rank)
    c = lambda v: v.float().to(DEV).contiguous()
    loss = torch.zeros(1, device=DEV)
    dls = torch.zeros(1, device=DEV)
    dui, dut = torch.empty(n, E, device=DEV), torch.empty(n, E, device=DEV)
    ws = torch.empty(ops.head_workspace_bytes(n, N, E) // 4, device=DEV)
    ops.head_fwd_bwd(c(ui), c(ut), c(ui_all), c(ut_all), c(t.reshape(1)), n, N, E, rank, 1.0, loss, dui, dut, dls, ws)
    assert abs(loss.item() - loss_ref.item()) < 1e-5 * max(1.0, abs(loss_ref.item()))
    assert rel(dui, dui_ref) < 2e-5 and rel(dut, dut_ref) < 2e-5
    assert abs(dls.item() - dt_ref.item()) < 2e-5 * max(1.0, abs(dt_ref.item()))


def test_adamw_and_sumsq_match_torch():
    ops = _ops()
    torch.manual_seed(9)
    n = 64 * 37
    p0 = torch.randn(n, device=DEV)
    flags = (torch.arange(n // 64, device=DEV) % 2).to(torch.uint8)
    ref_p = p0.clone().double()
    pa = torch.nn.Parameter(ref_p.clone())
    mask = flags.bool().repeat_interleave(64)
    opt = torch.optim.AdamW([{"params": [pa], "weight_decay": 0.0}], lr=5e-4, betas=(0.9, 0.98), eps=1e-6)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    pb = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    hyper = torch.empty(3, device=DEV)
    for step in range(1, 4):
        g = torch.randn(n, device=DEV) * (30.0 if step == 2 else 0.01)
        ss = torch.zeros(1, device=DEV)
        ops.sumsq(g, n, ss)
        assert rel(ss, (g.double() ** 2).sum().reshape(1)) < 1e-5
        hyper.copy_(torch.tensor([5e-4, 1 - 0.9 ** step, 1 - 0.98 ** step]))
        ops.adamw(p, g, m, v, pb, flags, n, ss, hyper, 1.0, 20.0, 0.9, 0.98, 1e-6, 0.2)
        # torch reference: clip_grad_norm_(20) then AdamW, decay only where flagged
        gd = g.double()
        coef = min(1.0, 20.0 / (gd.norm().item() + 1e-6))
        pa.grad = gd * coef
        with torch.no_grad():
            pa.data[mask] *= (1 - 5e-4 * 0.2)
        opt.step()
        assert rel(p, pa.data) < 1e-5, step
        assert torch.equal(pb, p.to(torch.bfloat16))
