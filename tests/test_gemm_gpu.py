"""GPU parity of the two GEMM engines (mc_gemm_bf16_tc, mc_gemm_f32_simt) through the C ABI.

A floating-point kernel: the checker here is a plain torch fp32/fp64 matmul of the same (bf16-rounded)
operands on the GPU.  Tolerances: fp32 outputs of the tensor-core engine 2e-4 relative to the
row scale (bf16 products are exact in fp32; only accumulation order differs), bf16 outputs 1e-2,
SIMT engine 1e-5.  Shapes cover every operand-major combination, ragged M/N/K (the 50/77/197-token
cases of SURVEY 7.3-1), batching, K-over-batch reductions, split-K and each epilogue branch.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from clip_mixer_b200 import ops
    return ops


def _store(logical, major, pad, dtype):
    """logical: [batch, R, K] float32.  Returns (tensor_in_memory, ld, batch_stride)."""
    b, R, K = logical.shape
    if major == 0:  # K contiguous
        ld = K + pad
        mem = torch.zeros(b, R, ld, device=logical.device, dtype=dtype)
        mem[:, :, :K] = logical.to(dtype)
        return mem, ld, R * ld
    ld = R + pad
    mem = torch.zeros(b, K, ld, device=logical.device, dtype=dtype)
    mem[:, :, :R] = logical.transpose(1, 2).to(dtype)
    return mem, ld, K * ld


def _gelu(z):
    return z * torch.sigmoid(1.702 * z)


def _gelu_grad(z):
    s = torch.sigmoid(1.702 * z)
    return s * (1 + 1.702 * z * (1 - s))


def scale_hint(K, batch, k_spans):
    return math.sqrt(K * (batch if k_spans else 1)) + 1.0


def run_case(engine, M, N, K, batch=1, a_major=0, b_major=0, a_shared=False, b_shared=False, k_spans=False,
             c_bf16=False, bias_mode=0, act=0, zout=False, residual=False, accumulate=False, split_k=1,
             row_remap=0, pad_a=0, pad_b=0, seed=0, rowsum=False, c_trans=False):
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(seed)
    opd = torch.bfloat16 if engine == "tc" else torch.float32
    zdt = torch.float16 if engine == "tc" else torch.float32
    ab = 1 if a_shared else batch
    bb = 1 if b_shared else batch
    A_log = torch.randn(ab, M, K, generator=g).to(dev)
    B_log = torch.randn(bb, N, K, generator=g).to(dev)
    A_mem, lda, a_bs = _store(A_log, a_major, pad_a, opd)
    B_mem, ldb, b_bs = _store(B_log, b_major, pad_b, opd)
    if a_shared:
        a_bs = 0
    if b_shared:
        b_bs = 0
    Ar = A_log.to(opd).double().expand(batch, M, K)
    Br = B_log.to(opd).double().expand(batch, N, K)
    acc = torch.einsum("bmk,bnk->bmn", Ar, Br)
    if k_spans:
        acc = acc.sum(0, keepdim=True)
    ob = acc.shape[0]
    Mout = M + (M // row_remap + 1 if row_remap else 0)
    x = acc.clone()
    bias = None
    if bias_mode == 1:
        bias = torch.randn(N, generator=g).to(dev)
        x = x + bias.double()[None, None, :]
    elif bias_mode == 2:
        bias = torch.randn(M, generator=g).to(dev)
        x = x + bias.double()[None, :, None]
    z_expected = x.clone()
    zin = None
    if act == 1:
        x = _gelu(x)
    elif act == 2:
        zin_log = torch.randn(ob, M, N, generator=g).to(dev).to(zdt)
        x = x * _gelu_grad(zin_log.double())
    R = None
    if residual:
        R_log = torch.randn(ob, M, N, generator=g).to(dev)
        x = x + R_log.double()

    def place(t, dtype, fill=0.0):  # logical [ob, M, N] -> memory [ob, Mout, N] honouring row_remap
        mem = torch.full((ob, Mout, N), fill, device=dev, dtype=dtype)
        if row_remap:
            rows = torch.arange(M, device=dev)
            mem[:, rows + rows // row_remap + 1, :] = t.to(dtype)
        else:
            mem[:] = t.to(dtype)
        return mem

    if act == 2:
        zin = place(zin_log, zdt)
    if residual:
        R = place(R_log, torch.float32)
    cdt = torch.bfloat16 if c_bf16 else torch.float32
    C0 = torch.randn(ob, Mout, N, generator=g).to(dev).to(cdt) if (accumulate or split_k > 1) else \
        torch.full((ob, Mout, N), 7.0, device=dev, dtype=cdt)
    Cm = C0.clone()
    Z = torch.full((ob, Mout, N), 7.0, device=dev, dtype=zdt) if zout else None
    rs = torch.full((M,), 3.0, device=dev) if rowsum else None
    if c_trans:   # C and R live as [ob, N, M] in memory
        Ct = torch.full((ob, N, M), 7.0, device=dev)
        Rt = R.transpose(1, 2).contiguous() if R is not None else None
        ops.gemm(engine, M, N, K, batch, A_mem, a_major, lda, a_bs, B_mem, b_major, ldb, b_bs, Ct, M, N * M, bias=bias,
                 bias_mode=bias_mode, R=Rt, ldr=M, r_bs=N * M, c_transposed=True)
        torch.cuda.synchronize()
        err = ((Ct.transpose(1, 2).double() - x).abs().max() / scale_hint(K, batch, k_spans)).item()
        assert err <= (2e-4 if engine == "tc" else 2e-5), f"transposed C mismatch {err:.3e}"
        return
    ops.gemm(engine, M, N, K, batch, A_mem, a_major, lda, a_bs, B_mem, b_major, ldb, b_bs, Cm, N, Mout * N,
             k_spans_batch=k_spans, accumulate=accumulate, split_k=split_k, row_remap=row_remap, bias=bias,
             bias_mode=bias_mode, zout=Z, ldz=N, z_bs=Mout * N, zin=zin, ldzin=N, zin_bs=Mout * N, act=act, R=R,
             ldr=N, r_bs=Mout * N, rowsum_out=rs)
    torch.cuda.synchronize()
    if rowsum:
        exp_rs = 3.0 + x.sum(dim=(0, 2))
        rerr = ((rs.double() - exp_rs).abs().max() / (scale_hint(K, batch, k_spans) * math.sqrt(N * x.shape[0]))).item()
        assert rerr <= (3e-3 if engine == "tc" else 2e-5), f"rowsum mismatch {rerr:.3e}"
    expected = place(x, torch.float64, fill=7.0)
    if accumulate or split_k > 1:
        expected = C0.double() + place(x, torch.float64, fill=0.0)
    scale = math.sqrt(K * (batch if k_spans else 1)) + 1.0
    tol = (2e-2 if c_bf16 else (3e-3 if act else 2e-4)) if engine == "tc" else 2e-5
    err = ((Cm.double() - expected).abs().max() / scale).item()
    detail = ""
    if not (err <= tol):
        bad = (Cm.double() - expected).abs() / scale > tol
        idx = bad.nonzero()
        detail = f" bad={int(bad.sum())}/{bad.numel()} first={idx[:4].tolist()} last={idx[-2:].tolist()}"
    assert err <= tol, f"C mismatch: max err/scale {err:.3e} > {tol}{detail}"
    if zout:
        ze = place(z_expected, torch.float64, fill=7.0)
        zerr = ((Z.double() - ze).abs().max() / scale).item()
        assert zerr <= (3e-3 if engine == "tc" else 2e-5), f"zout mismatch {zerr:.3e}"


ENGINES = ["simt", "tc"]


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_tn_aligned(engine):
    run_case(engine, 256, 256, 128)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_majors(engine, a_major, b_major):
    run_case(engine, 256, 384, 192, a_major=a_major, b_major=b_major)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_ragged(engine, a_major, b_major):
    # M, N, K none a multiple of the tile; leading dims padded to the 16-byte rule
    run_case(engine, 200, 72, 50, a_major=a_major, b_major=b_major, pad_a=6 if a_major == 0 else 0,
             pad_b=6 if b_major == 0 else 0)


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_lin3_like(engine):
    # bias along N + pre-activation store + QuickGELU, act-dtype output (channel-mix lin3)
    run_case(engine, 400, 512, 128, bias_mode=1, act=1, zout=True, c_bf16=(engine == "tc"))


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_lin4_like(engine):
    run_case(engine, 400, 128, 512, bias_mode=1, residual=True)


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_dgrad_gelu_bwd(engine):
    run_case(engine, 400, 512, 128, b_major=1, act=2, c_bf16=(engine == "tc"))


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_wgrad_accumulate(engine):
    run_case(engine, 512, 128, 400, a_major=1, b_major=1, accumulate=True)


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_wgrad_splitk(engine):
    run_case(engine, 256, 128, 4096, a_major=1, b_major=1, accumulate=True, split_k=4)


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_wgrad_auto_split(engine):
    run_case(engine, 200, 50, 768, batch=16, k_spans=True, accumulate=True, split_k=0)


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_token_mix_lin1(engine):
    # Z1[b] = W1[4P,P] @ U[b][P,D] + b1 (bias along M), P=50 (weights padded to ld 56), D=768
    run_case(engine, 200, 768, 50, batch=5, a_shared=True, b_major=1, bias_mode=2, act=1, zout=True, pad_a=6,
             c_bf16=(engine == "tc"))


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_token_mix_lin2(engine):
    run_case(engine, 77, 512, 308, batch=3, a_shared=True, b_major=1, bias_mode=2, residual=True, pad_a=4)


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_token_mix_dgrad(engine):
    # dZ1[b] = (W2^T @ dY[b]) * g'(Z1[b]): A is the MN-major view of W2 [P, 4P]
    run_case(engine, 200, 256, 50, batch=4, a_shared=True, a_major=1, b_major=1, act=2, c_bf16=(engine == "tc"))


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_token_mix_dgrad_fused_bias_grad(engine):
    # db1[h] += sum_{b,d} dZ1[b,h,d] rides on the dZ1 GEMM (rowsum_out); 2 and 3 m-tiles, ragged N
    run_case(engine, 200, 256, 50, batch=4, a_shared=True, a_major=1, b_major=1, act=2, c_bf16=(engine == "tc"), rowsum=True)
    run_case(engine, 308, 200, 77, batch=3, a_shared=True, a_major=1, b_major=1, act=2, c_bf16=(engine == "tc"), rowsum=True,
             pad_a=4)
    run_case(engine, 788, 96, 197, batch=2, a_shared=True, a_major=1, b_major=1, act=2, c_bf16=(engine == "tc"), rowsum=True,
             pad_a=4)


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_token_mix_d_as_m(engine):
    # lin2 with D as M: out[d, p] = sum_h H1[h, d] W2[p, h] + b2[p] + x[p, d], stored transposed ([P, D] slab)
    run_case(engine, 768, 50, 200, batch=5, a_major=1, b_shared=True, bias_mode=1, residual=True, c_trans=True)
    run_case(engine, 512, 77, 308, batch=3, a_major=1, b_shared=True, bias_mode=1, residual=True, c_trans=True, pad_b=4)
    # dU = W1^T dZ1 with D as M: both operands MN-major, plain transposed store; ragged D
    run_case(engine, 200, 17, 68, batch=4, a_major=1, b_major=1, b_shared=True, c_trans=True, pad_b=7)


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_token_mix_wgrad(engine):
    # dW1[4P,P] += sum_b dZ1[b] @ U[b]^T : reduction over (batch, D)
    run_case(engine, 200, 50, 256, batch=7, k_spans=True, accumulate=True)


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_patch_embed_row_remap(engine):
    # rows of patches land behind a class-token row: dest = m + m/49 + 1
    run_case(engine, 98, 64, 96, row_remap=49)


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_tiny_dims(engine):
    run_case(engine, 4, 32, 64, b_major=1)
    run_case(engine, 17, 64, 16, batch=3, a_shared=True, b_major=1, bias_mode=2)


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_many_tiles_persistent(engine):
    # more tiles than SMs: exercises the persistent loop, both TMEM accumulator stages and phase flips
    run_case(engine, 128 * 20, 256 * 10, 320, bias_mode=1)


def test_gemm_tc_b32_shapes():
    # the real channel-mix shapes at a reduced batch (image B32: D=768, text: 512)
    run_case("tc", 50 * 32, 3072, 768, bias_mode=1, act=1, zout=True, c_bf16=True)
    run_case("tc", 50 * 32, 768, 3072, bias_mode=1, residual=True)
    run_case("tc", 77 * 32, 2048, 512, bias_mode=1, act=1, zout=True, c_bf16=True)
    run_case("tc", 3072, 768, 50 * 32, a_major=1, b_major=1, accumulate=True, split_k=0)


def test_gemm_rejects_bad_args():
    ops = _ops()
    from clip_mixer_b200._lib import MixerClipError
    dev = torch.device("cuda:0")
    a = torch.zeros(16, 50, device=dev, dtype=torch.bfloat16)   # ld 50: 100-byte pitch violates the TMA rule
    b = torch.zeros(16, 50, device=dev, dtype=torch.bfloat16)
    c = torch.zeros(16, 16, device=dev)
    with pytest.raises(MixerClipError):
        ops.gemm("tc", 16, 16, 50, 1, a, 0, 50, 0, b, 0, 50, 0, c, 16, 0)
    with pytest.raises(MixerClipError):
        ops.gemm("tc", 16, 16, 50, 1, a.float(), 0, 50, 0, b, 0, 50, 0, c, 16, 0)


def test_gemm_tc_recompute_pair():
    """C = (A B^T) * QuickGELU'(A2 B2^T + bias2[m]) with the pre-activation recomputed by a second operand pair
    (token-mixing dZ1 without a saved Z1), plus the fused row sums; compared with the same math in fp64."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    for (M, N, K, batch, ldw) in ((200, 768, 50, 5, 56), (308, 512, 77, 3, 80), (68, 64, 17, 4, 24)):
        W2t = torch.randn(K, M, generator=g)                   # A as the MN-major view of W2 [P, 4P]
        dY = torch.randn(batch, K, N, generator=g)             # B MN-major
        W1 = torch.randn(M, K, generator=g)                    # A2 K-major, padded pitch
        U = torch.randn(batch, K, N, generator=g)              # B2 MN-major
        b1 = torch.randn(M, generator=g)
        ldm = (M + 7) // 8 * 8
        A = torch.zeros(K, ldm, dtype=torch.bfloat16, device=dev)
        A[:, :M] = W2t.to(dev)
        A2 = torch.zeros(M, ldw, dtype=torch.bfloat16, device=dev)
        A2[:, :K] = W1.to(dev)
        Bm, B2 = dY.to(dev).to(torch.bfloat16), U.to(dev).to(torch.bfloat16)
        C = torch.empty(batch, M, N, device=dev, dtype=torch.bfloat16)
        rs = torch.zeros(M, device=dev)
        ops.gemm("tc", M, N, K, batch, A, 1, ldm, 0, Bm, 1, N, K * N, C, N, M * N, act=ops.ACT_GELU_BWD, rowsum_out=rs,
                 A2=A2, a2_major=0, lda2=ldw, a2_bs=0, B2=B2, b2_major=1, ldb2=N, b2_bs=K * N, bias2=b1.to(dev))
        torch.cuda.synchronize()
        acc = torch.einsum("km,bkn->bmn", A[:, :M].double(), Bm.double())
        z = torch.einsum("mk,bkn->bmn", A2[:, :K].double(), B2.double()) + b1.to(dev).double()[None, :, None]
        ref = acc * _gelu_grad(z)
        scale = math.sqrt(K) + 1
        assert ((C.double() - ref).abs().max() / scale).item() <= 2e-2
        assert ((rs.double() - ref.sum((0, 2))).abs().max() / (scale * math.sqrt(N * batch))).item() <= 3e-3


@pytest.mark.parametrize("sm_limit", [0, 48, 80, 84, 64])
def test_gemm_tc_streamed_epilogues_on_a_share_of_the_sms(sm_limit):
    """Regression (round 2): the residual (lin4) and GELU-backward (dZ2) epilogues re-arm the TMA-loaded input slot right
    after reading it; without waiting for the ld.shared results, the next chunk's data could land first (16-byte pieces
    of wrong residual / pre-activation, sporadic, far more frequent when the grid covers a share of the SMs as in the
    two-stream training step).  Shapes of a 64-sample step (odd numbers of 128-row tiles), several SM shares."""
    ops = _ops()
    ops.set_sm_limit(sm_limit)
    try:
        run_case("tc", 50 * 64, 768, 3072, bias_mode=1, residual=True)                       # image lin4
        run_case("tc", 50 * 64, 3072, 768, b_major=1, act=2, c_bf16=True)                    # image dZ2
        run_case("tc", 77 * 64, 512, 2048, bias_mode=1, residual=True)                       # text lin4
        run_case("tc", 77 * 64, 2048, 512, b_major=1, act=2, c_bf16=True)                    # text dZ2
        run_case("tc", 50 * 64, 3072, 768, bias_mode=1, act=1, zout=True, c_bf16=True)       # image lin3
        run_case("tc", 768, 3072, 50 * 64, a_major=1, b_major=1, accumulate=True, split_k=0)  # dW4
    finally:
        ops.set_sm_limit(0)


def test_gemm_tc_residual_epilogue_row_statistics():
    """mc_gemm_params.rowstat_out: the lin4 epilogue leaves (sum, sum of squares) of every output row for the LayerNorm
    that the next block's fused token-mixing kernel runs in its prologue."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(3)
    M, N, K = 50 * 24, 768, 512
    A = torch.randn(M, K, generator=g).to(dev).to(torch.bfloat16)
    Bw = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(dev)
    R = torch.randn(M, N, generator=g).to(dev)
    C = torch.empty(M, N, device=dev)
    stat = torch.zeros(M, 2, device=dev)
    ops.gemm("tc", M, N, K, 1, A, 0, K, 0, Bw, 0, K, 0, C, N, 0, bias=bias, bias_mode=ops.BIAS_N, R=R, ldr=N, rowstat_out=stat)
    torch.cuda.synchronize()
    ref = A.double() @ Bw.double().t() + bias.double() + R.double()
    assert float((C.double() - ref).abs().max()) <= 2e-3
    assert float((stat[:, 0].double() - C.double().sum(1)).abs().max()) <= 1e-3 * float(C.double().abs().sum(1).max())
    assert float((stat[:, 1].double() - (C.double() ** 2).sum(1)).abs().max()) <= 1e-5 * float((C.double() ** 2).sum(1).max())
