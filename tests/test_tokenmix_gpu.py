"""GPU parity of the fused token-mixing kernels (mc_token_mix_fwd / _dgrad / _wgrad) through the C ABI.

Reference math: MixerBlock.token_mix, training/clip/model.py:206-208,216,220-222, restated on the same
bf16-rounded operands in fp64 torch (a floating-point kernel: the checker is a plain torch reference of the
same op).  Tolerances (written per assert): the outputs carry one bf16 rounding of the hidden activation
(2^-9 relative per element) through a K = 4P contraction, so 4e-3 of the output scale for fwd / dgrad and
1e-2 of the gradient's L2 norm for the weight gradients (K = B*D contraction of bf16-rounded factors).
Shapes: the two production towers (P=50/D=768, P=77/D=512), the smallest legal tile (P=5, D=128) and a
P that is a multiple of 16 (no token padding).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(3, 50, 768), (2, 77, 512), (4, 5, 128), (2, 64, 256), (300, 50, 256)]


def _ops():
    from clip_mixer_b200 import ops
    return ops


def _gelu(z):
    return z * torch.sigmoid(1.702 * z)


def _gelu_grad(z):
    s = torch.sigmoid(1.702 * z)
    return s * (1 + 1.702 * z * (1 - s))


def _setup(B, P, D, seed=0):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(seed)
    H = 4 * P
    ld1, ld2 = (P + 7) // 8 * 8, (H + 7) // 8 * 8
    w1 = torch.zeros(H, ld1)
    w1[:, :P] = torch.randn(H, P, generator=g) / P ** 0.5
    w2 = torch.zeros(P, ld2)
    w2[:, :H] = torch.randn(P, H, generator=g) / H ** 0.5
    # pad elements deliberately non-zero: the kernels must ignore them
    w1[:, P:] = 7.0
    w2[:, H:] = -5.0
    t = dict(
        u=torch.randn(B, P, D, generator=g).to(dev).to(torch.bfloat16),
        x=torch.randn(B, P, D, generator=g).to(dev),
        dy=torch.randn(B, P, D, generator=g).to(dev).to(torch.bfloat16),
        w1=w1.to(dev).to(torch.bfloat16), w2=w2.to(dev).to(torch.bfloat16),
        b1=torch.randn(H, generator=g).to(dev) * 0.5, b2=torch.randn(P, generator=g).to(dev) * 0.5,
        ld1=ld1, ld2=ld2, H=H)
    return t


def _reference(t, P):
    H = t["H"]
    W1 = t["w1"][:, :P].double()          # [H, P]
    W2 = t["w2"][:, :H].double()          # [P, H]
    U = t["u"].double()                   # [B, P, D]
    Z = torch.einsum("jp,bpd->bjd", W1, U) + t["b1"].double()[None, :, None]
    Hh = _gelu(Z)
    Hb = Hh.to(torch.bfloat16).double()   # the kernel feeds the second GEMM with bf16
    Y = t["x"].double() + torch.einsum("pj,bjd->bpd", W2, Hb) + t["b2"].double()[None, :, None]
    dY = t["dy"].double()
    dH = torch.einsum("pj,bpd->bjd", W2, dY)
    dZ = (dH * _gelu_grad(Z))
    dZb = dZ.to(torch.bfloat16).double()
    dU = torch.einsum("jp,bjd->bpd", W1, dZb)
    gW2 = torch.einsum("bpd,bjd->pj", dY, Hb)
    gW1 = torch.einsum("bjd,bpd->jp", dZb, U)
    gb1 = dZb.sum(dim=(0, 2))
    return dict(Y=Y, Hb=Hb, dZb=dZb, dU=dU, gW1=gW1, gW2=gW2, gb1=gb1)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("B,P,D", SHAPES)
def test_token_mix_fwd(B, P, D):
    ops = _ops()
    assert ops.token_mix_supported(P, D)
    t = _setup(B, P, D)
    ref = _reference(t, P)
    y = torch.full_like(t["x"], float("nan"))
    ops.token_mix_fwd(B, P, D, t["u"], t["x"], y, t["w1"], t["ld1"], t["b1"], t["w2"], t["ld2"], t["b2"])
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    delta = (y.double() - ref["Y"]).abs().max().item()
    scale = (ref["Y"] - t["x"].double()).abs().max().item()
    assert delta <= 4e-3 * scale + 1e-5, (delta, scale)
    assert _rel(y.double() - t["x"].double(), ref["Y"] - t["x"].double()) <= 4e-3


@pytest.mark.parametrize("B,P,D", SHAPES)
def test_token_mix_fwd_with_layernorm_in_the_prologue(B, P, D):
    """model.py:216 folded into :220-222: the kernel normalises the fp32 block input itself from per-row (sum, sum of
    squares) - what the producing lin4 GEMM leaves in rowstat_out - and also emits u (bf16) and the row statistics."""
    ops = _ops()
    t = _setup(B, P, D)
    dev = t["x"].device
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn(B, P, D, generator=g) * 1.7 + 0.4).to(dev)        # non-zero mean: exercises sumsq/D - mean^2
    gamma, beta = (torch.rand(D, generator=g) + 0.5).to(dev), (torch.randn(D, generator=g) * 0.3).to(dev)
    xd = x.double()
    sums = torch.stack([x.sum(-1), (x * x).sum(-1)], dim=-1).reshape(B * P, 2).contiguous()     # fp32, like the GEMM epilogue
    mean_ref = xd.mean(-1)
    rstd_ref = 1.0 / torch.sqrt(xd.var(-1, unbiased=False) + 1e-5)
    u_ref = ((xd - mean_ref[..., None]) * rstd_ref[..., None] * gamma.double() + beta.double())
    t2 = dict(t)
    t2["u"], t2["x"] = u_ref.to(torch.bfloat16), x
    ref = _reference(t2, P)
    y = torch.full_like(x, float("nan"))
    u_out = torch.full((B, P, D), float("nan"), device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.empty(B * P, device=dev), torch.empty(B * P, device=dev)
    ops.token_mix_fwd(B, P, D, None, x, y, t["w1"], t["ld1"], t["b1"], t["w2"], t["ld2"], t["b2"],
                      ln=dict(sums=sums, gamma=gamma, beta=beta, u_out=u_out, mean=mean, rstd=rstd))
    torch.cuda.synchronize()
    assert torch.isfinite(y).all() and torch.isfinite(u_out.float()).all()
    assert _rel(mean, mean_ref.reshape(-1)) <= 1e-5 and _rel(rstd, rstd_ref.reshape(-1)) <= 1e-4
    assert _rel(u_out, u_ref) <= 4e-3                                   # bf16 rounding of u
    delta = (y.double() - ref["Y"]).abs().max().item()
    scale = (ref["Y"] - xd).abs().max().item()
    assert delta <= 1.2e-2 * scale + 1e-5, (delta, scale)              # u may round to the neighbouring bf16 value
    assert _rel(y.double() - xd, ref["Y"] - xd) <= 6e-3


@pytest.mark.parametrize("B,P,D", SHAPES)
def test_token_mix_dgrad(B, P, D):
    ops = _ops()
    t = _setup(B, P, D, seed=1)
    ref = _reference(t, P)
    du = torch.full_like(t["x"], float("nan"))
    ops.token_mix_dgrad(B, P, D, t["u"], t["dy"], du, t["w1"], t["ld1"], t["b1"], t["w2"], t["ld2"])
    torch.cuda.synchronize()
    assert torch.isfinite(du).all()
    assert _rel(du, ref["dU"]) <= 4e-3
    assert (du.double() - ref["dU"]).abs().max().item() <= 6e-3 * ref["dU"].abs().max().item()
    # deterministic: a second launch gives the same bits
    du2 = torch.empty_like(du)
    ops.token_mix_dgrad(B, P, D, t["u"], t["dy"], du2, t["w1"], t["ld1"], t["b1"], t["w2"], t["ld2"])
    torch.cuda.synchronize()
    assert torch.equal(du, du2)


@pytest.mark.parametrize("B,P,D", SHAPES)
def test_token_mix_wgrad(B, P, D):
    ops = _ops()
    t = _setup(B, P, D, seed=2)
    ref = _reference(t, P)
    H = t["H"]
    dev = t["x"].device
    ldg1, ldg2 = t["ld1"], t["ld2"]
    g = torch.Generator(device="cpu").manual_seed(9)
    gw1_0 = torch.randn(H, ldg1, generator=g).to(dev)
    gw2_0 = torch.randn(P, ldg2, generator=g).to(dev)
    gb1_0 = torch.randn(H, generator=g).to(dev)
    gw1, gw2, gb1 = gw1_0.clone(), gw2_0.clone(), gb1_0.clone()
    ops.token_mix_wgrad(B, P, D, t["u"], t["dy"], t["w1"], t["ld1"], t["b1"], t["w2"], t["ld2"], gw1, ldg1, gw2, ldg2, gb1)
    torch.cuda.synchronize()
    d1 = (gw1 - gw1_0)[:, :P].double()
    d2 = (gw2 - gw2_0)[:, :H].double()
    db = (gb1 - gb1_0).double()
    assert _rel(d1, ref["gW1"]) <= 1e-2, _rel(d1, ref["gW1"])
    assert _rel(d2, ref["gW2"]) <= 1e-2, _rel(d2, ref["gW2"])
    assert _rel(db, ref["gb1"]) <= 1e-2, _rel(db, ref["gb1"])
    # pad columns of the gradient buffers are never touched (they must stay zero in the flat gradient store)
    assert torch.equal(gw1[:, P:], gw1_0[:, P:]) and torch.equal(gw2[:, H:], gw2_0[:, H:])


def test_token_mix_unsupported_shape_raises():
    ops = _ops()
    from clip_mixer_b200._lib import MixerClipError
    assert not ops.token_mix_supported(197, 768)
    assert not ops.token_mix_supported(50, 96)
    t = _setup(1, 50, 128)
    y = torch.empty_like(t["x"])
    with pytest.raises(MixerClipError):
        ops.token_mix_fwd(1, 197, 128, t["u"], t["x"], y, t["w1"], t["ld1"], t["b1"], t["w2"], t["ld2"], t["b2"])
    with pytest.raises(MixerClipError):     # x and y must not alias (the residual is read through the read-only path)
        ops.token_mix_fwd(1, 50, 128, t["u"], t["x"], t["x"], t["w1"], t["ld1"], t["b1"], t["w2"], t["ld2"], t["b2"])
