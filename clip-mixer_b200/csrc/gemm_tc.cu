// tcgen05 / TMEM / TMA batched GEMM engine with fused epilogue (sm_100a).
//
// One persistent, warp-specialised kernel serves every tensor-core GEMM of the Mixer-CLIP step
// (call sites k1, k4-k11, k16 of SURVEY.md 2.4; reference: nn.Linear lin1..lin4 + QuickGELU +
// residual, training/clip/model.py:206-222; conv1 as im2col GEMM :258,272; projections :288,424;
// and the dgrad / wgrad GEMMs autograd derives from them, training/training.py:170).
//
//   warps 0-11: epilogue       (tcgen05.ld 32x32b -> bias / QuickGELU / QuickGELU' / residual -> swizzled smem -> TMA
//                               store; three warps per TMEM lane quarter)
//   warp 12   : TMEM allocator (512 columns = 2 accumulator stages of up to 256 columns)
//   warp 14   : TMA producer   (cp.async.bulk.tensor.3d -> 128B-swizzled smem ring, mbarrier tx; elected lane)
//   warp 13   : second TMA producer (B operand) when prod2
//   warp 15   : MMA issuer     (elected lane, tcgen05.mma kind::f16, 128 / 256 x BN x 16, fp32 accum in TMEM)
//
// Operand delivery through the smem ring paces these GEMMs (~42 B/clk/SM measured through this pipeline; the memory
// system itself lands 114-125 B/clk/SM from L2, tools/ubench/l2_ingest.cu), so the fewer operand bytes a CTA needs per
// k-block the better: the kernel runs as CTA pairs (tcgen05 cta_group::2) computing a 256 x BN tile, each CTA stages its
// own 128 A rows and HALF of the B tile (32 KB instead of 48 KB per CTA per k-block).
//
// Operand layouts are described, not materialised: a K-major operand is one TMA box of
// [rows x 64] bf16 per stage; an MN-major operand (the "transposed" view: token-mixing reads the
// [P x D] activation in place, wgrad reads activations as [K=tokens x M]) is ceil(rows/64) boxes of
// [64 k-rows x 64] and the UMMA descriptor carries the MN-major canonical layout
// (LBO = 8192 B between 64-wide groups, SBO = 1024 B between 8-row k groups).
// Out-of-bounds parts of a box are zero-filled by TMA, so ragged M/N/K (50, 77, 197, 200, 308 ...)
// need no padding copies of the activations; only rows whose byte pitch is not a multiple of 16 B
// (token-mix weights) are re-packed by mc_cast_pad.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace mc {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements per k-block = 128 B = one swizzle row
constexpr int kEpiWarps = 12;                    // three per TMEM lane quarter
constexpr int kThreads = 128 + kEpiWarps * 32;  // 512
// Warp roles.  The scheduler of an SM sub-partition favours its highest warp id and a polling warp keeps taking issue
// slots, so the two warps everything else waits for (TMA producer, MMA issuer) sit above the twelve epilogue warps,
// and the epilogue warps back off with nanosleep while they wait for an accumulator (same finding as in tokenmix.cu).
constexpr int kAllocWarp = kEpiWarps, kProdWarpB = kEpiWarps + 1, kProdWarp = kEpiWarps + 2, kMmaWarp = kEpiWarps + 3;
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages
constexpr uint32_t kABytes = BM * BK * 2;  // 16384
constexpr uint32_t kGroupBytes = 64 * BK * 2;  // one [64 x 64] bf16 box = 8192

struct GemmTcArgs {
    int M, N, K;
    int out_batch, kb_per_batch, kb_total, k_spans_batch;
    int a_mn, b_mn, a_batched, b_batched;
    int BN, stages, stage_bytes, b_tx_bytes;
    int tiles_m, tiles_n, split_k, num_tiles;
    int cluster, tiles_m_eff, b_box_rows;   // cluster: CTAs per cluster along M (1 or 2)
    int tma_epi;                            // outputs leave through smem staging + TMA store
    int two_cta;                            // tcgen05 cta_group::2: the pair computes a 256 x BN tile, B split in halves
    int prod2;                              // second producer warp issues the B operand
    int exp_flags;                          // MC_GEMM_EXP bits (default 3 since round 2): 1 plain remote arrive, 2 lean MMA issuer loop
    int l2pf_a, l2pf_b;                     // k-blocks of L2 prefetch distance for the A / B operand (0 = off)
    int epi_bufs, epi_alt_off;              // output staging buffers per epilogue warp (1 or 2) and the byte offset of the second
    int zdepth, epi_warp_bytes;             // lookahead of the epilogue's TMA-loaded inputs (chunks), staging bytes per epilogue warp
    // MC_GEMM_DEBUG_SKIP bits (timing experiments, wrong results): 1 no TMA loads, 2 no MMAs, 4 no epilogue.  Without the
    // loads a stage's full barrier no longer depends on the second producer warp, so that warp can be lapped by two
    // phases and the watchdog traps the launch (seen under ncu and with the residual epilogue): use with MC_GEMM_PROD2=0
    // and treat a trap as "rerun", never as a product failure.
    int dbg_skip;
    int epi_smem_off;                       // byte offset of the epilogue staging area from tiles_base
    // epilogue
    void* C;
    int c_bf16;
    long long ldc, c_bs;
    int accumulate, atomic, row_remap, vec_ok;
    const float* bias;
    int bias_mode;
    __half* zout;  // pre-activations are private to this engine: fp16 (11-bit mantissa), see DESIGN.md
    long long ldz, z_bs;
    const __half* zin;
    long long ldzin, zin_bs;
    int act;
    const float* R;
    long long ldr, r_bs;
    float* rowsum_out;   // optional: rowsum_out[m] += sum over (batch, n) of the stored value
    float* rowstat_out;  // optional (residual epilogue): rowstat_out[2m] += sum_n x, [2m+1] += sum_n x^2 of the stored value
    int c_transposed;    // C / R element (m, n) at n*ld + m
    // second operand pair (recompute of the pre-activation): acc2 lands 128 TMEM columns after acc
    int dual, a2_mn, b2_mn, a2_batched, b2_batched, pair_bytes;
    const float* bias2;
};

struct TileCoord {
    int tm, tn, b, kb_begin, kb_end;
};

// t indexes work items of a cluster; the CTA of rank r in the cluster takes m-tile tm_eff * cluster + r
__device__ __forceinline__ TileCoord decode_tile(const GemmTcArgs& g, int t, int cta_rank) {
    // n-tiles vary fastest: the CTAs working at any moment share a few A row-blocks and sweep the (small,
    // L2-resident) B operand, so the big activation operand is streamed from DRAM once
    TileCoord c;
    c.tn = t % g.tiles_n;
    t /= g.tiles_n;
    c.tm = (t % g.tiles_m_eff) * g.cluster + cta_rank;
    t /= g.tiles_m_eff;
    c.b = t % g.out_batch;
    const int ks = t / g.out_batch;
    c.kb_begin = int((long long)ks * g.kb_total / g.split_k);
    c.kb_end = int((long long)(ks + 1) * g.kb_total / g.split_k);
    return c;
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    lo = fminf(fmaxf(lo, -65504.f), 65504.f);
    hi = fminf(fmaxf(hi, -65504.f), 65504.f);
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }

// ---- epilogue on 8 consecutive columns held in registers ----------------------------------------
__device__ __forceinline__ void epilogue8(const GemmTcArgs& g, float (&x)[8], float bias_m, long long crow, int b,
                                          int n, int ncols, bool vec, float& rsum) {
    // bias
    if (g.bias_mode == MC_BIAS_N) {
        if (vec) {
            const float4 b0 = *reinterpret_cast<const float4*>(g.bias + n);
            const float4 b1 = *reinterpret_cast<const float4*>(g.bias + n + 4);
            x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
            x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < ncols) x[j] += g.bias[n + j];
        }
    } else if (g.bias_mode == MC_BIAS_M) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += bias_m;
    }
    // pre-activation store
    if (g.zout != nullptr) {
        __half* zp = g.zout + (long long)b * g.z_bs + crow * g.ldz + n;
        if (vec) {
            uint4 pk;
            pk.x = pack_h2(x[0], x[1]); pk.y = pack_h2(x[2], x[3]);
            pk.z = pack_h2(x[4], x[5]); pk.w = pack_h2(x[6], x[7]);
            *reinterpret_cast<uint4*>(zp) = pk;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < ncols) zp[j] = __float2half_rn(fminf(fmaxf(x[j], -65504.f), 65504.f));
        }
    }
    // activation
    if (g.act == MC_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = quick_gelu_fast(x[j]);
    } else if (g.act == MC_ACT_GELU_BWD) {
        const __half* zp = g.zin + (long long)b * g.zin_bs + crow * g.ldzin + n;
        if (vec) {
            const uint4 pk = *reinterpret_cast<const uint4*>(zp);
            const float2 z0 = unpack_h2(pk.x), z1 = unpack_h2(pk.y), z2 = unpack_h2(pk.z), z3 = unpack_h2(pk.w);
            x[0] *= quick_gelu_grad_fast(z0.x); x[1] *= quick_gelu_grad_fast(z0.y);
            x[2] *= quick_gelu_grad_fast(z1.x); x[3] *= quick_gelu_grad_fast(z1.y);
            x[4] *= quick_gelu_grad_fast(z2.x); x[5] *= quick_gelu_grad_fast(z2.y);
            x[6] *= quick_gelu_grad_fast(z3.x); x[7] *= quick_gelu_grad_fast(z3.y);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < ncols) x[j] *= quick_gelu_grad_fast(__half2float(zp[j]));
        }
    }
    // residual
    if (g.R != nullptr) {
        const float* rp = g.R + (long long)b * g.r_bs + crow * g.ldr + n;
        if (vec) {
            const float4 r0 = *reinterpret_cast<const float4*>(rp);
            const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
            x[0] += r0.x; x[1] += r0.y; x[2] += r0.z; x[3] += r0.w;
            x[4] += r1.x; x[5] += r1.y; x[6] += r1.z; x[7] += r1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < ncols) x[j] += rp[j];
        }
    }
    if (g.rowsum_out != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < ncols) rsum += x[j];
    }
    // store
    if (g.c_bf16) {
        __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(g.C) + (long long)b * g.c_bs + crow * g.ldc + n;
        if (vec) {
            uint4 pk;
            pk.x = pack_bf16x2(x[0], x[1]); pk.y = pack_bf16x2(x[2], x[3]);
            pk.z = pack_bf16x2(x[4], x[5]); pk.w = pack_bf16x2(x[6], x[7]);
            *reinterpret_cast<uint4*>(cp) = pk;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < ncols) cp[j] = __float2bfloat16_rn(x[j]);
        }
    } else {
        float* cp = reinterpret_cast<float*>(g.C) + (long long)b * g.c_bs + crow * g.ldc + n;
        if (g.atomic) {
            if (vec) {
                atomicAdd(reinterpret_cast<float4*>(cp), make_float4(x[0], x[1], x[2], x[3]));
                atomicAdd(reinterpret_cast<float4*>(cp + 4), make_float4(x[4], x[5], x[6], x[7]));
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < ncols) atomicAdd(cp + j, x[j]);
            }
        } else if (vec) {
            float4 o0 = make_float4(x[0], x[1], x[2], x[3]), o1 = make_float4(x[4], x[5], x[6], x[7]);
            if (g.accumulate) {
                const float4 c0 = *reinterpret_cast<const float4*>(cp);
                const float4 c1 = *reinterpret_cast<const float4*>(cp + 4);
                o0.x += c0.x; o0.y += c0.y; o0.z += c0.z; o0.w += c0.w;
                o1.x += c1.x; o1.y += c1.y; o1.z += c1.z; o1.w += c1.w;
            }
            *reinterpret_cast<float4*>(cp) = o0;
            *reinterpret_cast<float4*>(cp + 4) = o1;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < ncols) cp[j] = g.accumulate ? cp[j] + x[j] : x[j];
        }
    }
}

// ---- 256-bit global accesses -------------------------------------------------------------------------
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void ldg256f(const float* p, float (&r)[8]) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void stg256f(float* p, const float (&r)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]),
                 "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
                 : "memory");
}
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

enum { EPI_GENERIC = 0, EPI_ACT_FWD = 1, EPI_RESID = 2, EPI_ACT_BWD = 3, EPI_PLAIN = 4, EPI_TRANS = 5, EPI_ACT_BWD_DUAL = 6 };
constexpr int kDualAccOffset = 128;  // TMEM columns between acc and acc2 (dual mode: BN <= 128)

// QuickGELU on the tanh unit: x*sigmoid(1.702x) = hx + hx*tanh(0.851x), hx = x/2
__device__ __forceinline__ float gelu_t(float x) {
    const float hx = 0.5f * x;
    return fmaf(hx, tanh_approx(0.851f * x), hx);
}
// derivative for the backward epilogues: sigmoid from ex2 + rcp (both ~1 ulp) instead of tanh.approx (2^-11):
// the error of g'(z) multiplies every gradient that flows through the block
__device__ __forceinline__ float gelu_grad_t(float z) {
    float e, s;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.4554669595930157f * z));   // exp(-1.702 z)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(1.0f + e));
    const float w = kGeluA * z;
    return fmaf(fmaf(-w, s, w), s, s);  // s + w*s*(1-s)
}

// Packed (2 x fp32 per instruction, FFMA2 / FMUL2 / FADD2) versions for the staged epilogues, which are
// instruction-issue bound on the token-mixing shapes (profiles/r1d): QuickGELU = hx + hx*tanh(0.851x).
__device__ __forceinline__ float2 gelu2(float2 x) {
    const float2 a = __fmul2_rn(x, make_float2(0.851f, 0.851f));
    const float2 t = make_float2(tanh_approx(a.x), tanh_approx(a.y));
    const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    return __ffma2_rn(hx, t, hx);
}
#ifndef MC_GELU_GRAD_TANH
#define MC_GELU_GRAD_TANH 1
#endif
__device__ __forceinline__ float2 gelu_grad2(float2 z) {
#if MC_GELU_GRAD_TANH
    // sigmoid(1.702 z) = 0.5 + 0.5 tanh(0.851 z): ONE MUFU op per element instead of two (ex2 + rcp).  The GELU-backward
    // epilogue is the slowest of the engine (64 us on its own for the B/32 dZ2 tile set, MC_GEMM_DEBUG_SKIP=3) and two
    // MUFU ops per element are 4096 clocks per 128 x 256 tile - the whole MMA time of a K = 512 tile.  tanh.approx is
    // good to 2^-11 relative: 1e-4 absolute on g', against the 2e-3 of the bf16 rounding of the result.
    const float2 a = __fmul2_rn(z, make_float2(0.851f, 0.851f));
    const float2 t = make_float2(tanh_approx(a.x), tanh_approx(a.y));
    const float2 s = __ffma2_rn(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
#else
    const float2 a = __fmul2_rn(z, make_float2(-2.4554669595930157f, -2.4554669595930157f));   // -1.702*log2(e)*z
    float ex, ey, sx, sy;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(a.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(a.y));
    const float2 den = __fadd2_rn(make_float2(ex, ey), make_float2(1.0f, 1.0f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sx) : "f"(den.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sy) : "f"(den.y));
    const float2 s = make_float2(sx, sy);
#endif
    const float2 w = __fmul2_rn(z, make_float2(kGeluA, kGeluA));
    const float2 nw = __fmul2_rn(z, make_float2(-kGeluA, -kGeluA));
    const float2 t1 = __ffma2_rn(nw, s, w);      // w (1 - s)
    return __ffma2_rn(t1, s, s);                 // s + w s (1 - s)
}

// One full 32-column chunk (all columns valid, 32-byte aligned rows), specialised per epilogue kind.
// Global operands of the chunk (residual / saved pre-activation) are requested BEFORE the TMEM load is
// waited for, so their latency overlaps it.
template <int EPI>
__device__ __forceinline__ void chunk32(const GemmTcArgs& g, uint32_t taddr, float bias_m, long long crow, int b, int n,
                                        bool ok, float& rsum) {
    // tcgen05.ld is warp-collective (.sync.aligned): EVERY lane executes it; only the global accesses are
    // predicated on the row being inside M.
    uint32_t v[32];
    if constexpr (EPI == EPI_RESID) {
        const float* rp = g.R + (long long)b * g.r_bs + crow * g.ldr + n;
        float r[4][8];
        if (ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) ldg256f(rp + 8 * j, r[j]);
        }
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        if (!ok) return;
        float* cp = reinterpret_cast<float*>(g.C) + (long long)b * g.c_bs + crow * g.ldc + n;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float o[8];
            if (g.bias_mode == MC_BIAS_N) {
                float bv[8];
                ldg256f(g.bias + n + 8 * j, bv);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v[8 * j + i]) + bv[i] + r[j][i];
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v[8 * j + i]) + bias_m + r[j][i];
            }
            stg256f(cp + 8 * j, o);
        }
    } else if constexpr (EPI == EPI_ACT_BWD) {
        const __half* zp = g.zin + (long long)b * g.zin_bs + crow * g.ldzin + n;
        uint32_t z[2][8];
        if (ok) {
            ldg256(zp, z[0]);
            ldg256(zp + 16, z[1]);
        }
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        if (!ok) return;
        __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(g.C) + (long long)b * g.c_bs + crow * g.ldc + n;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 zz = unpack_h2(z[j][i]);
                const float a = __uint_as_float(v[16 * j + 2 * i]) * gelu_grad_t(zz.x);
                const float c = __uint_as_float(v[16 * j + 2 * i + 1]) * gelu_grad_t(zz.y);
                rsum += a + c;
                o[i] = pack_bf16x2(a, c);
            }
            stg256(cp + 16 * j, o);
        }
    } else if constexpr (EPI == EPI_ACT_FWD) {
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        if (!ok) return;
        __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(g.C) + (long long)b * g.c_bs + crow * g.ldc + n;
        __half* zp = g.zout != nullptr ? g.zout + (long long)b * g.z_bs + crow * g.ldz + n : nullptr;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float x[16];
            if (g.bias_mode == MC_BIAS_N) {
                float bv[2][8];
                ldg256f(g.bias + n + 16 * j, bv[0]);
                ldg256f(g.bias + n + 16 * j + 8, bv[1]);
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[16 * j + i]) + bv[i >> 3][i & 7];
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[16 * j + i]) + bias_m;
            }
            if (zp != nullptr) {
                uint32_t zo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) zo[i] = pack_h2_sat(x[2 * i], x[2 * i + 1]);
                stg256(zp + 16 * j, zo);
            }
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(gelu_t(x[2 * i]), gelu_t(x[2 * i + 1]));
            stg256(cp + 16 * j, o);
        }
    } else {  // EPI_PLAIN: fp32 store / read-modify-write / atomic add, no bias, no activation
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        if (!ok) return;
        float* cp = reinterpret_cast<float*>(g.C) + (long long)b * g.c_bs + crow * g.ldc + n;
        if (g.atomic) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                atomicAdd(reinterpret_cast<float4*>(cp + 4 * j),
                          make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                      __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v[8 * j + i]);
                if (g.accumulate) {
                    float c0[8];
                    ldg256f(cp + 8 * j, c0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] += c0[i];
                }
                stg256f(cp + 8 * j, o);
            }
        }
    }
}

// Transposed fp32 output (D-as-M token-mixing GEMMs): the thread owns row m, so for every column the 32 lanes of
// the warp touch 32 consecutive addresses - each residual load and each store is one coalesced 128-byte line.
__device__ __forceinline__ void chunk32_trans(const GemmTcArgs& g, uint32_t taddr, float bias_m, int m, bool row_ok, int b,
                                              int n) {
    uint32_t v[32];
    float r[32];
    const int ncols = g.N - n < 32 ? g.N - n : 32;
    if (g.R != nullptr && row_ok) {
        const float* rp = g.R + (long long)b * g.r_bs + (long long)n * g.ldr + m;
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = j < ncols ? rp[(long long)j * g.ldr] : 0.f;
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0.f;
    }
    tmem_ld32(taddr, v);
    tmem_ld_wait();
    if (!row_ok) return;
    float* cp = reinterpret_cast<float*>(g.C) + (long long)b * g.c_bs + (long long)n * g.ldc + m;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j < ncols) {
            const float bb = g.bias_mode == MC_BIAS_N ? g.bias[n + j] : bias_m;
            cp[(long long)j * g.ldc] = __uint_as_float(v[j]) + bb + r[j];
        }
    }
}

// ---- TMA store side ------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// staging tile about to be rewritten: with two alternating output buffers per warp only the store issued TWO chunks ago
// must have finished reading shared memory, so the TMA store of chunk i drains while chunk i + 1 is computed
__device__ __forceinline__ void bulk_wait_stage(int bufs) {
    if (bufs > 1) bulk_wait_read1();
    else bulk_wait_read0();
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// The epilogue's TMA-loaded input slot is re-armed (overwritten by the next TMA load) right after it has been read.
// "Read" must mean the ld.shared results have ARRIVED: the loads are asynchronous, and with a provably converged warp
// the compiler drops the WARPSYNC of __syncwarp(), so nothing else sits between the LDS instructions and lane 0's
// UTMALDG.  A real instruction that consumes the last register of every load makes the warp wait on their scoreboards
// (round 2: sporadic 16-byte pieces of stale / next-chunk data in the residual and GELU-backward epilogues, 8-12 k wrong
// elements per GEMM when the grid covered a share of the SMs - found by the 256-sample parity test).
__device__ __forceinline__ void wait_for_loaded(uint32_t sink_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    // volatile store of a value derived from every load: ptxas cannot drop it (a plain xor chain with an unused result
    // is dead code to ptxas even inside asm volatile), and it cannot issue before the loads' scoreboards clear
    asm volatile("{\n\t.reg .b32 t;\n\txor.b32 t, %1, %2;\n\txor.b32 t, t, %3;\n\txor.b32 t, t, %4;\n\t"
                 "st.volatile.shared.b32 [%0], t;\n\t}"
                 ::"r"(sink_addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

constexpr int kChunkStride = 32 * (kEpiWarps / 4);   // columns between consecutive chunks of one epilogue warp
constexpr uint32_t kEpiWarpBytes = 4096;  // one 32x32 fp32 chunk, or a bf16 C chunk (2 KB) + an fp16 Z chunk (2 KB); +2 KB per extra Z slot

// One 32-column chunk leaving through shared memory and a TMA store: every lane owns one output row, writes it
// into a swizzled staging tile (conflict-free 16-byte stores), and one lane hands the [32 x 32] box to the TMA
// unit, which writes full lines to L2 and clips rows >= M / columns >= N by itself.  Compared with per-lane
// global stores (one 32-byte sector per lane = 32 L1 wavefronts per instruction) this removes ~90 % of the
// epilogue's L1 traffic, which ncu showed to be what kept the tensor pipe at 48 % (profiles/r1b_*).
template <int EPI>
__device__ __forceinline__ void chunk32_staged(const GemmTcArgs& g, const CUtensorMap* tmC, const CUtensorMap* tmZ,
                                               uint32_t taddr, uint32_t stage, uint32_t cstage, float bias_m, bool row_ok, long long crow,
                                               int m_base, int b, int n, int lane, uint32_t zbar, uint32_t zphase,
                                               uint32_t zoff, int pf_n, int pf_m, int pf_b, float& rsum,
                                               [[maybe_unused]] uint32_t sink_addr) {
    uint32_t v[32];
    [[maybe_unused]] const bool full = n + 32 <= g.N;
    if constexpr (EPI == EPI_ACT_BWD_DUAL) {
        // the pre-activation is RECOMPUTED by a second GEMM of the same tile (acc2, 128 TMEM columns further):
        // z = acc2 + bias2[m]; nothing is read from memory for it
        uint32_t w[32];
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + kDualAccOffset, w);
        tmem_ld_wait();
        const float2 b2 = make_float2(bias_m, bias_m);     // bias_m carries bias2[m] in this mode
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float2 z = __fadd2_rn(make_float2(__uint_as_float(w[2 * i]), __uint_as_float(w[2 * i + 1])), b2);
            const float2 r = __fmul2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), gelu_grad2(z));
            o[i] = pack_bf16x2(r.x, r.y);
            if (full) rsum += r.x + r.y;
            else rsum += (n + 2 * i < g.N ? r.x : 0.f) + (n + 2 * i + 1 < g.N ? r.y : 0.f);
        }
        if (lane == 0) bulk_wait_stage(g.epi_bufs);
        __syncwarp();
        const uint32_t rowp = cstage + lane * 64, sw = (lane >> 1) & 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) sts128(rowp + ((j ^ sw) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_3d(tmC, cstage, n, m_base, b);
            bulk_commit();
        }
    } else if constexpr (EPI == EPI_ACT_BWD) {
        // The saved pre-activation chunk (fp16 [32 x 32], swizzle-64B) was requested by TMA one or two chunks ahead
        // (across tile boundaries, see the epilogue loop) into slot `zoff` of the staging area; rows >= M / columns
        // >= N arrive as zeros.  The slot is handed to the next outstanding request (pf_*) as soon as it has been read.
        mbar_wait(zbar, zphase);
        const uint32_t rowp = stage + lane * 64, sw = (lane >> 1) & 3;
        uint32_t z[16];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(z[4 * j]), "=r"(z[4 * j + 1]), "=r"(z[4 * j + 2]), "=r"(z[4 * j + 3])
                         : "r"(rowp + zoff + ((j ^ sw) << 4)));
        wait_for_loaded(sink_addr, z[3], z[7], z[11], z[15]);
        __syncwarp();
        if (lane == 0 && pf_n >= 0) {
            mbar_arrive_expect_tx(zbar, 2048);
            tma_load_3d(stage + zoff, tmZ, zbar, pf_n, pf_m, pf_b);
        }
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        uint32_t o[16];
        if (g.rowsum_out == nullptr) {
            // channel-mixing dZ2 (the common case): no fused row sums.  The ncu source view counted 466 instructions per
            // chunk here against 227 in the forward epilogue, ~150 of them (FADD / ISETP / FSEL / VIADD) the row-sum
            // bookkeeping that only the token-mixing bias gradient of the GEMM schedule consumes.
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 gp = gelu_grad2(unpack_h2(z[i]));
                const float2 r = __fmul2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), gp);
                o[i] = pack_bf16x2(r.x, r.y);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 gp = gelu_grad2(unpack_h2(z[i]));
                const float2 r = __fmul2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), gp);
                o[i] = pack_bf16x2(r.x, r.y);
                if (full) rsum += r.x + r.y;                     // columns >= N hold garbage accumulators
                else rsum += (n + 2 * i < g.N ? r.x : 0.f) + (n + 2 * i + 1 < g.N ? r.y : 0.f);
            }
        }
        if (lane == 0) bulk_wait_stage(g.epi_bufs);
        __syncwarp();
        const uint32_t crowp = cstage + lane * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) sts128(crowp + ((j ^ sw) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_3d(tmC, cstage, n, m_base, b);
            bulk_commit();
        }
    } else if constexpr (EPI == EPI_ACT_FWD) {
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        float2 x[16];
        if (g.bias_mode == MC_BIAS_N) {
            if (full && g.vec_ok) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float bv[8];
                    ldg256f(g.bias + n + 8 * j, bv);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        x[4 * j + i] = __fadd2_rn(make_float2(__uint_as_float(v[8 * j + 2 * i]), __uint_as_float(v[8 * j + 2 * i + 1])),
                                                  make_float2(bv[2 * i], bv[2 * i + 1]));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    x[i] = make_float2(__uint_as_float(v[2 * i]) + (n + 2 * i < g.N ? g.bias[n + 2 * i] : 0.f),
                                       __uint_as_float(v[2 * i + 1]) + (n + 2 * i + 1 < g.N ? g.bias[n + 2 * i + 1] : 0.f));
            }
        } else {
            const float2 bm = make_float2(bias_m, bias_m);
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fadd2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), bm);
        }
        if (lane == 0) bulk_wait_stage(g.epi_bufs);
        __syncwarp();
        const uint32_t rowp = cstage + lane * 64, sw = (lane >> 1) & 3;
        if (g.zout != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                sts128(rowp + 2048 + ((j ^ sw) << 4), pack_h2_sat(x[4 * j].x, x[4 * j].y), pack_h2_sat(x[4 * j + 1].x, x[4 * j + 1].y),
                       pack_h2_sat(x[4 * j + 2].x, x[4 * j + 2].y), pack_h2_sat(x[4 * j + 3].x, x[4 * j + 3].y));
        }
        uint32_t h[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float2 y = gelu2(x[i]);
            h[i] = pack_bf16x2(y.x, y.y);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) sts128(rowp + ((j ^ sw) << 4), h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_3d(tmC, cstage, n, m_base, b);
            if (g.zout != nullptr) tma_store_3d(tmZ, cstage + 2048, n, m_base, b);
            bulk_commit();
        }
    } else if constexpr (EPI == EPI_RESID) {
        // The fp32 residual chunk ([32 x 32], swizzle-128B) was requested by TMA one chunk ahead; the result
        // leaves through direct 256-bit stores (write-only, nothing waits on them).
        mbar_wait(zbar, zphase);
        const uint32_t rowp = stage + lane * 128, sw = lane & 7;
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r[4 * j]), "=r"(r[4 * j + 1]), "=r"(r[4 * j + 2]), "=r"(r[4 * j + 3])
                         : "r"(rowp + ((j ^ sw) << 4)));
        wait_for_loaded(sink_addr, r[3] ^ r[7], r[11] ^ r[15], r[19] ^ r[23], r[27] ^ r[31]);
        __syncwarp();
        if (lane == 0 && pf_n >= 0) {
            mbar_arrive_expect_tx(zbar, 4096);
            tma_load_3d(stage, tmZ, zbar, pf_n, pf_m, pf_b);
        }
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        if (!row_ok) return;
        float* cp = reinterpret_cast<float*>(g.C) + (long long)b * g.c_bs + crow * g.ldc + n;
        float st1 = 0.f, st2 = 0.f;      // row statistics of the stored values (rowstat_out)
        if (full && g.vec_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float o[8];
                if (g.bias_mode == MC_BIAS_N) {
                    float bv[8];
                    ldg256f(g.bias + n + 8 * j, bv);
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v[8 * j + i]) + bv[i] + __uint_as_float(r[8 * j + i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v[8 * j + i]) + bias_m + __uint_as_float(r[8 * j + i]);
                }
                stg256f(cp + 8 * j, o);
                if (g.rowstat_out != nullptr) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        st1 += o[i];
                        st2 = fmaf(o[i], o[i], st2);
                    }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (n + i < g.N) {
                    const float bb = g.bias_mode == MC_BIAS_N ? g.bias[n + i] : bias_m;
                    const float o = __uint_as_float(v[i]) + bb + __uint_as_float(r[i]);
                    cp[i] = o;
                    st1 += o;
                    st2 = fmaf(o, o, st2);
                }
            }
        }
        if (g.rowstat_out != nullptr) {
            atomicAdd(g.rowstat_out + 2 * crow, st1);
            atomicAdd(g.rowstat_out + 2 * crow + 1, st2);
        }
    } else {  // EPI_PLAIN: fp32 tile, plain store or reduce-add (accumulate / split-K)
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        if (lane == 0) bulk_wait_stage(g.epi_bufs);
        __syncwarp();
        const uint32_t rowp = cstage + lane * 128, sw = lane & 7;
#pragma unroll
        for (int j = 0; j < 8; ++j) sts128(rowp + ((j ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            if (g.accumulate || g.atomic) tma_reduce_add_3d(tmC, cstage, n, m_base, b);
            else tma_store_3d(tmC, cstage, n, m_base, b);
            bulk_commit();
        }
    }
}

// L2 prefetch of one TMA box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
}

// ---- cta_group::2 (CTA pair) primitives -----------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
// default-semantics arrive on a barrier of another CTA of the cluster (no cluster-scope release fence).  The
// release.cluster form below costs ~1 k clocks (tools/ubench/l2_ingest cluster rows; MEMBAR.ALL.CTA + ERRBAR + CGAERRBAR
// in SASS, 8-11 % of the non-leader epilogue warps' ncu samples).  The accumulator hand-back only has to order the
// tcgen05.ld reads, which tcgen05.fence::before_thread_sync does.  MC_GEMM_EXP bit 1 (default on: dZ2 875 -> 934 TFLOP/s).
__device__ __forceinline__ void mbar_arrive_cluster_plain(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on a barrier that may live in the peer (leader) CTA
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_ss_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int EPI, bool TWO>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmZ,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const GemmTcArgs g) {
    extern __shared__ uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ __align__(8) uint64_t zin_bar[kEpiWarps][2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ uint32_t epi_sink[kEpiWarps];      // see wait_for_loaded()

    // A role index the compiler can prove warp-uniform turns the role branches into uniform control flow, so the
    // producer / issuer / epilogue loops keep their counters, barrier addresses and descriptors in uniform registers:
    // R2UR moves in this kernel 744 -> 37, 5800 -> 4870 SASS instructions, the issuer's k-loop becomes ~35
    // uniform-datapath instructions (cuobjdump).  Measured on B200 (profiles/r2_first_experiments.txt): lin4 957 -> 1039,
    // dW3 1171 -> 1259 TFLOP/s, whole step 15.18 -> 14.70 ms.  -DMC_DIVERGENT_WARP_IDX restores the plain index for A/B.
#ifndef MC_DIVERGENT_WARP_IDX
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
#else
    const int warp = threadIdx.x >> 5;
#endif
    const int lane = threadIdx.x & 31;
    const uint32_t tiles_base = (smem_u32(dyn_smem) + 1023u) & ~1023u;
    const int csize = g.cluster;
    const int cta_rank = csize > 1 ? (int)cluster_ctarank() : 0;
    const int work0 = blockIdx.x / csize, work_stride = gridDim.x / csize;
    // cta_group::2 pair mode is a separate instantiation: a kernel that contains pair instructions can only be
    // launched with an even cluster size
    constexpr bool two = TWO;              // rank 0 is the leader and issues every MMA
    const bool leader = cta_rank == 0;

    if (warp == kProdWarp && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (g.tma_epi) {
            if (EPI != EPI_RESID) tma_prefetch_desc(&tmC);
            if (g.zout != nullptr || g.zin != nullptr || g.R != nullptr) tma_prefetch_desc(&tmZ);
        }
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            // multicast mode: released by the MMA warp of every CTA of the cluster; pair mode: one multicast
            // commit of the leader's MMA thread reaches both CTAs
            mbar_init(smem_u32(&empty_bar[s]), two ? 1 : csize);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tfull_bar[s]), 1);
            // pair mode: the leader's accumulator-free barrier collects the epilogue warps of BOTH CTAs
            mbar_init(smem_u32(&tempty_bar[s]), two ? 2 * kEpiWarps : kEpiWarps);
        }
        for (int s = 0; s < kEpiWarps; ++s) {
            mbar_init(smem_u32(&zin_bar[s][0]), 1);
            mbar_init(smem_u32(&zin_bar[s][1]), 1);
        }
        fence_mbar_init();
    }
    if (warp == kAllocWarp) {
        if constexpr (TWO) {
            tmem_alloc_2cta(smem_u32(&tmem_base_smem), kTmemCols);
            tmem_relinquish_2cta();
        } else {
            tmem_alloc(smem_u32(&tmem_base_smem), kTmemCols);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (csize > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    // everything above touched only shared / tensor memory and kernel parameters: with programmatic dependent launch it
    // overlaps the tail of the preceding kernel; global memory is first touched below
    MC_PDL_PROLOGUE();

    if (warp == kProdWarp || (warp == kProdWarpB && g.prod2)) {
        // ===================== TMA producer(s) =====================
        // Like the MMA issuer, the whole warp runs the loop on warp-uniform values and one elected lane issues: a
        // single-lane `if (lane == 0)` body makes the compiler wrap every UTMALDG in an ELECT / R2UR.BROADCAST /
        // BRA.U.ANY waterfall, and the ncu source view of the r1s2 kernels showed the producer busy issuing all the time
        // (almost never waiting for a free stage) while the MMA warp waited for data 43 % of its time - ~1200 clocks per
        // k-block against 512 of MMA work.  With prod2 the (otherwise idle) second producer warp issues the B operand
        // while this one issues A; both watch the same empty barrier, the A warp posts the expected byte count (bytes
        // that complete before the expect_tx only make the transaction count dip below zero, the phase cannot flip
        // before the arrival).
        const bool do_a = warp == kProdWarp;
        const bool do_b = g.prod2 ? warp == kProdWarpB : true;
        uint32_t is_issuer;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_issuer));
        const bool issuer = is_issuer != 0;
        uint32_t stage = 0, phase = 0;
        const uint16_t mc_mask = (uint16_t)((1u << csize) - 1u);
        // Rolling L2 prefetch (MC_GEMM_L2PF): a second cursor walks the same (tile, k-block) sequence `dist` k-blocks ahead
        // of the loads and asks L2 for the boxes of the operands that stream from DRAM (activations).  The ring covers
        // ~2.5 k clocks of MMA work, less than a loaded DRAM round trip, so without it the issuer waits on `full`.
        const int pf_dist = do_a ? g.l2pf_a : g.l2pf_b;
        int pf_t = work0, pf_kb = 0;
        bool pf_valid = false;
        TileCoord pfc{};
        auto pf_tile = [&]() {
            pf_valid = pf_dist > 0 && pf_t < g.num_tiles;
            if (pf_valid) {
                pfc = decode_tile(g, pf_t, cta_rank);
                pf_kb = pfc.kb_begin;
            }
        };
        auto pf_step = [&]() {
            if (!pf_valid) return;
            int pb = pfc.b, pk = pf_kb * BK;
            if (g.k_spans_batch) {
                pb = pf_kb / g.kb_per_batch;
                pk = (pf_kb - pb * g.kb_per_batch) * BK;
            }
            if (issuer) {
                if (do_a && g.l2pf_a) {
                    const int pm0 = pfc.tm * BM, pba = g.a_batched ? pb : 0;
                    if (g.a_mn) {
                        tma_prefetch_l2_3d(&tmA, pm0, pk, pba);
                        tma_prefetch_l2_3d(&tmA, pm0 + 64, pk, pba);
                    } else {
                        tma_prefetch_l2_3d(&tmA, pk, pm0, pba);
                    }
                }
                if (do_b && g.l2pf_b) {
                    const int pbb = g.b_batched ? pb : 0;
                    const int cols = TWO ? g.BN / 2 : g.BN;
                    const int pn0 = pfc.tn * g.BN + (TWO ? cta_rank * (g.BN / 2) : 0);
                    if (g.b_mn) {
                        for (int j = 0; j * 64 < cols; ++j) tma_prefetch_l2_3d(&tmB, pn0 + j * 64, pk, pbb);
                    } else if (TWO || csize == 1) {
                        tma_prefetch_l2_3d(&tmB, pk, pn0, pbb);
                    }
                }
            }
            if (++pf_kb >= pfc.kb_end) {
                pf_t += work_stride;
                pf_tile();
            }
        };
        pf_tile();
        for (int d = 0; d < pf_dist; ++d) pf_step();
        for (int t = work0; t < g.num_tiles; t += work_stride) {
            const TileCoord tc = decode_tile(g, t, cta_rank);
            const int m0 = tc.tm * BM, n0 = tc.tn * g.BN;
            for (int kb = tc.kb_begin; kb < tc.kb_end; ++kb) {
                int bb = tc.b, kk = kb * BK;
                if (g.k_spans_batch) {
                    bb = kb / g.kb_per_batch;
                    kk = (kb - bb * g.kb_per_batch) * BK;
                }
                pf_step();
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
                const uint32_t bar = smem_u32(&full_bar[stage]);
                const uint32_t a_dst = tiles_base + stage * g.stage_bytes;
                const uint32_t b_dst = a_dst + kABytes;
                const int ba = g.a_batched ? bb : 0, bbt = g.b_batched ? bb : 0;
                if (g.dbg_skip & 1) {
                    if (issuer && do_a && (!TWO || leader)) mbar_arrive(bar);
                } else if (issuer) {
                    if constexpr (TWO) {
                        // pair mode: both CTAs' loads complete on the LEADER's barrier, which expects the bytes of both
                        const uint32_t lbar = leader ? bar : mapa_u32(bar, 0);
                        if (do_a) {
                            if (leader) mbar_arrive_expect_tx(bar, 2u * (kABytes + g.b_tx_bytes));
                            if (g.a_mn) {
                                tma_load_3d_2sm(a_dst, &tmA, lbar, m0, kk, ba);
                                tma_load_3d_2sm(a_dst + kGroupBytes, &tmA, lbar, m0 + 64, kk, ba);
                            } else {
                                tma_load_3d_2sm(a_dst, &tmA, lbar, kk, m0, ba);
                            }
                        }
                        if (do_b) {
                            const int nh = n0 + cta_rank * (g.BN / 2);   // this CTA supplies half of the B tile's columns
                            if (g.b_mn) {
                                for (int j = 0; j * 64 < g.BN / 2; ++j)
                                    tma_load_3d_2sm(b_dst + j * kGroupBytes, &tmB, lbar, nh + j * 64, kk, bbt);
                            } else {
                                tma_load_3d_2sm(b_dst, &tmB, lbar, kk, nh, bbt);
                            }
                        }
                    } else {
                        if (do_a) {
                            mbar_arrive_expect_tx(bar, (g.dual ? 2u : 1u) * (kABytes + g.b_tx_bytes));
                            if (g.a_mn) {
                                tma_load_3d(a_dst, &tmA, bar, m0, kk, ba);
                                tma_load_3d(a_dst + kGroupBytes, &tmA, bar, m0 + 64, kk, ba);
                            } else {
                                tma_load_3d(a_dst, &tmA, bar, kk, m0, ba);
                            }
                            if (g.dual) {   // second operand pair of the same tile (single-CTA mode only)
                                const uint32_t a2_dst = a_dst + g.pair_bytes, b2_dst = a2_dst + kABytes;
                                const int ba2 = g.a2_batched ? bb : 0, bb2 = g.b2_batched ? bb : 0;
                                if (g.a2_mn) {
                                    tma_load_3d(a2_dst, &tmA2, bar, m0, kk, ba2);
                                    tma_load_3d(a2_dst + kGroupBytes, &tmA2, bar, m0 + 64, kk, ba2);
                                } else {
                                    tma_load_3d(a2_dst, &tmA2, bar, kk, m0, ba2);
                                }
                                if (g.b2_mn) {
                                    for (int j = 0; j * 64 < g.BN; ++j)
                                        tma_load_3d(b2_dst + j * kGroupBytes, &tmB2, bar, n0 + j * 64, kk, bb2);
                                } else {
                                    tma_load_3d(b2_dst, &tmB2, bar, kk, n0, bb2);
                                }
                            }
                        }
                        if (do_b) {
                            if (csize == 1) {
                                if (g.b_mn) {
                                    for (int j = 0; j * 64 < g.BN; ++j)
                                        tma_load_3d(b_dst + j * kGroupBytes, &tmB, bar, n0 + j * 64, kk, bbt);
                                } else {
                                    tma_load_3d(b_dst, &tmB, bar, kk, n0, bbt);
                                }
                            } else {
                                // this CTA fetches its share of the B tile and multicasts it to the whole cluster
                                if (g.b_mn) {
                                    for (int j = cta_rank; j * 64 < g.BN; j += csize)
                                        tma_load_3d_mc(b_dst + j * kGroupBytes, &tmB, bar, n0 + j * 64, kk, bbt, mc_mask);
                                } else {
                                    const int r0 = cta_rank * g.b_box_rows;
                                    tma_load_3d_mc(b_dst + r0 * (BK * 2), &tmB, bar, kk, n0 + r0, bbt, mc_mask);
                                }
                            }
                        }
                    }
                }
                if (++stage == (uint32_t)g.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        // The whole warp runs the loop (all values are warp-uniform, so the descriptors live in uniform registers) and
        // one elected lane issues the four tcgen05.mma of a k-block back to back.  A single-lane `if (lane == 0)` body
        // costs ~137 clocks per MMA in R2UR / ELECT overhead (tools/ubench/mma_rate.cu) - more than the 128 clocks a
        // 128 x 256 x 16 MMA occupies the tensor pipe.
        if ((!two || leader) && (g.exp_flags & 2) && !g.dual) {
            // ---- lean issuer (MC_GEMM_EXP bit 2, default on) ----
            // tools/ubench/ring_handover runs this very pipeline at 515 clocks per k-block where the loop below needs
            // ~690 with the loads switched off; its SASS builds the descriptors with a handful of uniform-datapath
            // instructions, the loop below with a dozen R2UR moves of 64-bit templates kept in vector registers.  Here
            // the descriptor is assembled from 32-bit pieces inside the loop: low word = start address >> 4 | LBO field,
            // high word = SBO / version / swizzle constant.
            uint32_t is_issuer;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_issuer));
            const bool issuer = is_issuer != 0;
            const uint32_t idesc = make_idesc_bf16(two ? 2 * BM : BM, g.BN, g.a_mn, g.b_mn);
            constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);     // SBO 1024 B, version 1, 128B swizzle
            const uint32_t a_lbo = (g.a_mn ? (kGroupBytes >> 4) : 1u) << 16, b_lbo = (g.b_mn ? (kGroupBytes >> 4) : 1u) << 16;
            const uint32_t a_ks = g.a_mn ? 128u : 2u, b_ks = g.b_mn ? 128u : 2u;
            const uint16_t mc_mask = (uint16_t)((1u << csize) - 1u);
            const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
            const uint32_t nstages = (uint32_t)g.stages, stage_bytes = (uint32_t)g.stage_bytes;
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0, a_addr = tiles_base;
            for (int t = work0; t < g.num_tiles; t += work_stride) {
                const TileCoord tc = decode_tile(g, t, cta_rank);
                mbar_wait(smem_u32(&tempty_bar[as]), aphase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * kAccStride;
                uint32_t acc = 0u;
                for (int kb = tc.kb_begin; kb < tc.kb_end; ++kb) {
                    mbar_wait(full0 + 8u * stage, phase);
                    tc_fence_after();
                    const uint32_t a_lo = (a_addr >> 4) | a_lbo, b_lo = ((a_addr + kABytes) >> 4) | b_lbo;
                    if (issuer) {
#pragma unroll
                        for (uint32_t k = 0; k < BK / 16; ++k) {
                            const uint64_t ad = (uint64_t(kDescHi) << 32) | (a_lo + k * a_ks);
                            const uint64_t bd = (uint64_t(kDescHi) << 32) | (b_lo + k * b_ks);
                            if constexpr (TWO) umma_ss_2cta(d_tmem, ad, bd, idesc, k > 0 ? 1u : acc);
                            else umma_ss(d_tmem, ad, bd, idesc, k > 0 ? 1u : acc);
                        }
                        if constexpr (TWO) umma_commit_2cta_mc(empty0 + 8u * stage, mc_mask);
                        else if (csize == 1) umma_commit(empty0 + 8u * stage);
                        else umma_commit_mc(empty0 + 8u * stage, mc_mask);
                    }
                    acc = 1u;
                    a_addr += stage_bytes;
                    if (++stage == nstages) {
                        stage = 0;
                        phase ^= 1u;
                        a_addr = tiles_base;
                    }
                }
                if (issuer) {
                    if constexpr (TWO) umma_commit_2cta_mc(smem_u32(&tfull_bar[as]), mc_mask);
                    else umma_commit(smem_u32(&tfull_bar[as]));
                }
                as ^= 1u;
                if (as == 0) aphase ^= 1u;
            }
        } else if (!two || leader) {
            uint32_t is_issuer;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_issuer));
            const bool issuer = is_issuer != 0;
            const uint32_t idesc = make_idesc_bf16(two ? 2 * BM : BM, g.BN, g.a_mn, g.b_mn);
            const uint32_t idesc2 = make_idesc_bf16(BM, g.BN, g.a2_mn, g.b2_mn);
            // descriptor templates; the start address (>> 4) is added per instruction, a k-step of 16 advances it by
            // 32 B (K-major) or 2048 B (MN-major)
            const uint64_t a_tmpl = make_sdesc_sw128(0u, g.a_mn ? kGroupBytes : 16u, 1024u);
            const uint64_t b_tmpl = make_sdesc_sw128(0u, g.b_mn ? kGroupBytes : 16u, 1024u);
            const uint64_t a2_tmpl = make_sdesc_sw128(0u, g.a2_mn ? kGroupBytes : 16u, 1024u);
            const uint64_t b2_tmpl = make_sdesc_sw128(0u, g.b2_mn ? kGroupBytes : 16u, 1024u);
            const uint32_t a_ks = g.a_mn ? 128u : 2u, b_ks = g.b_mn ? 128u : 2u;
            const uint32_t a2_ks = g.a2_mn ? 128u : 2u, b2_ks = g.b2_mn ? 128u : 2u;
            const uint16_t mc_mask = (uint16_t)((1u << csize) - 1u);
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            for (int t = work0; t < g.num_tiles; t += work_stride) {
                const TileCoord tc = decode_tile(g, t, cta_rank);
                mbar_wait(smem_u32(&tempty_bar[as]), aphase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * kAccStride;
                for (int kb = tc.kb_begin; kb < tc.kb_end; ++kb) {
                    mbar_wait(smem_u32(&full_bar[stage]), phase);
                    tc_fence_after();
                    const uint32_t a_base = tiles_base + stage * g.stage_bytes;
                    const uint32_t b_base = a_base + kABytes;
                    const uint64_t ad = a_tmpl + (a_base >> 4), bd = b_tmpl + (b_base >> 4);
                    const uint32_t acc0 = kb > tc.kb_begin ? 1u : 0u;
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            if (g.dbg_skip & 2) break;
                            if constexpr (TWO) umma_ss_2cta(d_tmem, ad + k * a_ks, bd + k * b_ks, idesc, k > 0 ? 1u : acc0);
                            else umma_ss(d_tmem, ad + k * a_ks, bd + k * b_ks, idesc, k > 0 ? 1u : acc0);
                        }
                        if (!TWO && g.dual) {
                            const uint32_t a2_base = a_base + g.pair_bytes, b2_base = a2_base + kABytes;
                            const uint64_t ad2 = a2_tmpl + (a2_base >> 4), bd2 = b2_tmpl + (b2_base >> 4);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_ss(d_tmem + kDualAccOffset, ad2 + k * a2_ks, bd2 + k * b2_ks, idesc2, k > 0 ? 1u : acc0);
                        }
                        // the stage is shared through multicast / the pair: release it in every CTA of the cluster
                        if constexpr (TWO) umma_commit_2cta_mc(smem_u32(&empty_bar[stage]), mc_mask);
                        else if (csize == 1) umma_commit(smem_u32(&empty_bar[stage]));
                        else umma_commit_mc(smem_u32(&empty_bar[stage]), mc_mask);
                    }
                    if (++stage == (uint32_t)g.stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (issuer) {
                    if constexpr (TWO) umma_commit_2cta_mc(smem_u32(&tfull_bar[as]), mc_mask);
                    else umma_commit(smem_u32(&tfull_bar[as]));
                }
                as ^= 1u;
                if (as == 0) aphase ^= 1u;
            }
        }
    } else if (warp < kEpiWarps) {
        // ===================== epilogue =====================
        const int e = warp;
        const int q = e & 3;    // TMEM lane quarter == warp id % 4
        const int half = e >> 2;  // which 32-column chunks of every kChunkStride this warp drains (0 .. kEpiWarps/4-1)
        uint32_t as = 0, aphase = 0;
        const uint32_t stage_buf = tiles_base + g.epi_smem_off + e * g.epi_warp_bytes;
        // Stream of TMA-loaded epilogue inputs (saved pre-activation / residual chunks): the chunks this warp will drain
        // are known in advance (tile t, t + stride, ...; columns half*32, + kChunkStride, ...), so a cursor runs
        // g.zdepth chunks ahead of the consumer, across tile boundaries, and every slot is re-armed right after it has
        // been read.  Measured (profiles/r1s3_gemm_experiments.txt): neither two chunks of lookahead (MC_GEMM_ZDEPTH=2, costs a
        // ring stage) nor L2 promotion of these loads (MC_GEMM_ZPROMO) moves the GELU-backward epilogue (63 us on its own
        // for the B/32 dZ2 GEMM against 44 us for the forward epilogue), so the default stays at one chunk.
        constexpr bool kStream = EPI == EPI_ACT_BWD || EPI == EPI_RESID;
        const bool stream = kStream && g.tma_epi;
        int pf_t = work0, pf_c = half * 32, pf_n0 = 0, pf_m = 0, pf_b = 0;
        auto pf_settle = [&]() {   // move the cursor to the next (tile, chunk) that exists for this warp
            while (pf_t < g.num_tiles) {
                const TileCoord pc = decode_tile(g, pf_t, cta_rank);
                pf_n0 = pc.tn * g.BN; pf_m = pc.tm * BM + q * 32; pf_b = pc.b;
                if (pf_c < g.BN && pf_n0 + pf_c < g.N) return;
                pf_t += work_stride;
                pf_c = half * 32;
            }
        };
        uint32_t zslot = 0, zphase_bits = 0, cpar = 0;
        if (stream) {
            pf_settle();
            for (int d = 0; d < g.zdepth; ++d) {
                if (pf_t < g.num_tiles) {
                    if (lane == 0) {
                        const uint32_t zb = smem_u32(&zin_bar[e][d]);
                        mbar_arrive_expect_tx(zb, EPI == EPI_RESID ? 4096u : 2048u);
                        tma_load_3d(stage_buf + (EPI == EPI_RESID ? 0u : 2048u + 2048u * d), &tmZ, zb, pf_n0 + pf_c, pf_m, pf_b);
                    }
                    pf_c += kChunkStride;
                    pf_settle();
                }
            }
        }
        float rs_acc[2] = {0.f, 0.f};   // fused row sums, one slot per m-tile this CTA can meet (tiles_m_eff <= 2)
        int rs_row[2] = {-1, -1};
        for (int t = work0; t < g.num_tiles; t += work_stride) {
            const TileCoord tc = decode_tile(g, t, cta_rank);
            const int m = tc.tm * BM + q * 32 + lane;
            const int n0 = tc.tn * g.BN;
            const bool row_ok = m < g.M;
            const long long crow = g.row_remap > 0 ? (long long)m + m / g.row_remap + 1 : (long long)m;
            float bias_m = 0.f;
            if (g.bias_mode == MC_BIAS_M && row_ok) bias_m = g.bias[m];
            if (EPI == EPI_ACT_BWD_DUAL && g.bias2 != nullptr && row_ok) bias_m = g.bias2[m];
            mbar_wait_relaxed(smem_u32(&tfull_bar[as]), aphase, 64);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + as * kAccStride;
            float rsum = 0.f;
            for (int c = half * 32; c < g.BN; c += kChunkStride) {
                const int nb = n0 + c;
                if (nb >= g.N || (g.dbg_skip & 4)) break;  // warp-uniform
                const int rem = g.N - nb;
                if (EPI == EPI_TRANS) {
                    chunk32_trans(g, t_row + c, bias_m, m, row_ok, tc.b, nb);
                } else if (EPI != EPI_GENERIC && g.tma_epi) {
                    const bool pf_ok = kStream && pf_t < g.num_tiles;
                    const uint32_t cstage = stage_buf + (cpar ? (uint32_t)g.epi_alt_off : 0u);
                    if (g.epi_bufs > 1) cpar ^= 1u;
                    chunk32_staged<EPI>(g, &tmC, &tmZ, t_row + c, stage_buf, cstage, bias_m, row_ok, crow, tc.tm * BM + q * 32, tc.b,
                                        nb, lane, smem_u32(&zin_bar[e][zslot]), (zphase_bits >> zslot) & 1u,
                                        EPI == EPI_RESID ? 0u : 2048u + 2048u * zslot, pf_ok ? pf_n0 + pf_c : -1, pf_m, pf_b, rsum,
                                        smem_u32(&epi_sink[e]));
                    if (kStream) {
                        zphase_bits ^= 1u << zslot;
                        if (++zslot == (uint32_t)g.zdepth) zslot = 0;
                        if (pf_ok) {
                            pf_c += kChunkStride;
                            pf_settle();
                        }
                    }
                } else if (EPI != EPI_GENERIC && rem >= 32) {
                    chunk32<EPI>(g, t_row + c, bias_m, crow, tc.b, nb, row_ok, rsum);
                } else {
                    uint32_t v[32];
                    tmem_ld32(t_row + c, v);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) {
                            const int ncols = rem - j8 * 8;
                            if (ncols > 0) {
                                float x[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[j8 * 8 + j]);
                                epilogue8(g, x, bias_m, crow, tc.b, nb + j8 * 8, ncols, g.vec_ok && ncols >= 8, rsum);
                            }
                        }
                    }
                }
            }
            if (g.rowsum_out != nullptr && row_ok) {
                if (g.tiles_m_eff <= 2) {
                    const int slot = (tc.tm / g.cluster) & 1;
                    rs_acc[slot] += rsum;
                    rs_row[slot] = m;
                } else {
                    atomicAdd(g.rowsum_out + m, rsum);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (two && !leader) {
                    if (g.exp_flags & 1) mbar_arrive_cluster_plain(mapa_u32(smem_u32(&tempty_bar[as]), 0));
                    else mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));
                }
                else mbar_arrive(smem_u32(&tempty_bar[as]));
            }
            as ^= 1u;
            if (as == 0) aphase ^= 1u;
        }
        if (e == 0 && lane == 0) pdl_launch_dependents();   // this CTA's tiles are done: the next kernel's grid may be scheduled
        if (g.rowsum_out != nullptr) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (rs_row[k] >= 0) atomicAdd(g.rowsum_out + rs_row[k], rs_acc[k]);
        }
        if (g.tma_epi && EPI != EPI_RESID && lane == 0) bulk_wait_all();  // staging tiles must outlive their TMA stores
    }

    tc_fence_before();
    if (csize > 1) cluster_sync_all(); else __syncthreads();
    if (warp == kAllocWarp) {
        if constexpr (TWO) tmem_dealloc_2cta(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- host: tensor maps ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// operand with logical [batch][rows][K]; see header for major / ld / batch_stride.
int make_operand_map(CUtensorMap* map, const void* ptr, int major, int64_t rows, int64_t K, int64_t ld,
                     int64_t batch, int64_t batch_stride, int box_rows, const char* name) {
    EncodeTiledFn enc = get_encode_fn();
    MC_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    MC_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "gemm %s: base pointer must be 16-byte aligned", name);
    MC_CHECK((ld * 2) % 16 == 0, "gemm %s: leading dimension %lld (bf16) must be a multiple of 8 elements", name,
             (long long)ld);
    MC_CHECK(batch_stride == 0 || (batch_stride * 2) % 16 == 0, "gemm %s: batch stride must be a multiple of 8", name);
    const bool batched = batch_stride != 0 && batch > 1;
    cuuint64_t gdim[3], gstride[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    if (major == MC_MAJOR_K) {
        gdim[0] = (cuuint64_t)K;
        gdim[1] = (cuuint64_t)rows;
        box[0] = BK;
        box[1] = (cuuint32_t)box_rows;
    } else {
        gdim[0] = (cuuint64_t)rows;
        gdim[1] = (cuuint64_t)K;
        box[0] = 64;
        box[1] = BK;
    }
    gdim[2] = batched ? (cuuint64_t)batch : 1;
    box[2] = 1;
    gstride[0] = (cuuint64_t)ld * 2;
    gstride[1] = batched ? (cuuint64_t)batch_stride * 2 : gstride[0] * gdim[1];
    MC_CHECK(gdim[0] >= 1 && gdim[1] >= 1, "gemm %s: empty operand", name);
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MC_CHECK(r == CUDA_SUCCESS,
             "cuTensorMapEncodeTiled(%s) failed with %d (dims %llu,%llu,%llu strides %llu,%llu box %u,%u)", name,
             (int)r, (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
             (unsigned long long)gstride[0], (unsigned long long)gstride[1], box[0], box[1]);
    return MC_OK;
}

// Modelled cost (SM cycles) of running the problem with a given BN / split / cluster size.  A k-block of
// a 128 x BN tile costs max(tensor time 2*BN, L2->SM operand delivery at ~42 B/clk/SM); tiles are dealt to
// sms/cluster work slots in rounds (wave quantisation dominates at these sizes).
double model_cost(int64_t M, int64_t N, int64_t out_batch, int64_t kb_total, int BN, int split, int sms, int heavy_epi,
                  int cluster) {
    const int64_t tiles_m_eff = ceil_div(ceil_div(M, BM), cluster);
    const int64_t work = tiles_m_eff * ceil_div(N, BN) * out_batch * split;
    const int64_t rounds = ceil_div(work, sms / cluster);
    const double kb = double(kb_total) / split;
    const double bytes = (BM + double(BN) / cluster) * BK * 2;
    const double t_l2 = bytes / 38.0, t_mma = 2.0 * BN;   // operand delivery through the ring ~36-40 B/clk/SM (profiles/r1c sweep)
    const double epi = BN * (heavy_epi ? 14.0 : 6.0);
    double tile = kb * (t_mma > t_l2 ? t_mma : t_l2);
    if (epi > tile) tile = epi;  // epilogue of tile i overlaps the MMAs of tile i+1
    return rounds * (tile + 500.0) + epi + 2500.0;
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// output tensor [batch][rows][cols] (cols contiguous) as a TMA store target with a [32 x 32] box
int make_store_map(CUtensorMap* map, void* ptr, CUtensorMapDataType dt, int esz, int64_t cols, int64_t rows, int64_t batch,
                   int64_t ld, int64_t batch_stride, const char* name, int l2_promotion_bytes = 0) {
    EncodeTiledFn enc = get_encode_fn();
    MC_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    const bool batched = batch_stride != 0 && batch > 1;
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, batched ? (cuuint64_t)batch : 1};
    cuuint64_t gstride[2] = {(cuuint64_t)ld * esz, batched ? (cuuint64_t)batch_stride * esz : (cuuint64_t)ld * esz * rows};
    cuuint32_t box[3] = {32, 32, 1}, estr[3] = {1, 1, 1};
    // maps that are LOADED through (saved pre-activations: 64-byte row segments of lines whose other half another warp
    // asks for a moment later) let L2 fetch whole 128 / 256-byte blocks from DRAM
    const CUtensorMapL2promotion promo = l2_promotion_bytes >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                         : l2_promotion_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                                     : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    CUresult r = enc(map, dt, 3, ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     esz == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MC_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with %d", name, (int)r);
    return MC_OK;
}

template <int EPI, bool TWO>
int launch_gemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmZ,
                  const CUtensorMap& tmA2, const CUtensorMap& tmB2, const GemmTcArgs& g, int grid, size_t smem,
                  cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        MC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<EPI, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)g.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    MC_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<EPI, TWO>, tmA, tmB, tmC, tmZ, tmA2, tmB2, g));
    return MC_OK;
}

template <int EPI>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmZ,
                const CUtensorMap& tmA2, const CUtensorMap& tmB2, const GemmTcArgs& g, int grid, size_t smem,
                cudaStream_t stream) {
    return g.two_cta ? launch_gemm_t<EPI, true>(tmA, tmB, tmC, tmZ, tmA2, tmB2, g, grid, smem, stream)
                     : launch_gemm_t<EPI, false>(tmA, tmB, tmC, tmZ, tmA2, tmB2, g, grid, smem, stream);
}

}  // namespace

}  // namespace mc

using namespace mc;

extern "C" int mc_gemm_bf16_tc(const mc_gemm_params* p, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MC_CHECK(p != nullptr, "null params");
    MC_CHECK(p->M > 0 && p->N > 0 && p->K > 0 && p->batch > 0, "gemm: empty problem M=%lld N=%lld K=%lld batch=%lld",
             (long long)p->M, (long long)p->N, (long long)p->K, (long long)p->batch);
    MC_CHECK(p->M < (1ll << 31) && p->N < (1ll << 31) && p->K < (1ll << 31), "gemm: dimension too large");
    MC_CHECK(p->A && p->B && p->C, "gemm: null operand");
    MC_CHECK(p->c_dtype == MC_F32 || p->c_dtype == MC_BF16, "gemm: bad c_dtype");
    MC_CHECK(!(p->accumulate && p->c_dtype != MC_F32), "gemm: accumulate needs fp32 C");
    const bool dual = p->A2 != nullptr;
    MC_CHECK(p->act != MC_ACT_GELU_BWD || p->zin != nullptr || dual, "gemm: GELU_BWD needs zin (or the A2/B2 recompute pair)");
    MC_CHECK(!dual || (p->B2 != nullptr && p->act == MC_ACT_GELU_BWD && p->zin == nullptr && p->c_dtype == MC_BF16 &&
                       !p->k_spans_batch && !p->accumulate && p->row_remap == 0 && p->R == nullptr && p->zout == nullptr &&
                       p->bias_mode == MC_BIAS_NONE),
             "gemm: the A2/B2 pair is for C(bf16) = acc * QuickGELU'(acc2 + bias2[m]) only");
    MC_CHECK(p->bias_mode == MC_BIAS_NONE || p->bias != nullptr, "gemm: bias_mode set without bias");

    GemmTcArgs g{};
    g.M = (int)p->M; g.N = (int)p->N; g.K = (int)p->K;
    g.k_spans_batch = p->k_spans_batch ? 1 : 0;
    g.out_batch = g.k_spans_batch ? 1 : (int)p->batch;
    g.kb_per_batch = (int)ceil_div(p->K, BK);
    g.kb_total = g.kb_per_batch * (g.k_spans_batch ? (int)p->batch : 1);
    g.a_mn = p->a_major == MC_MAJOR_MN; g.b_mn = p->b_major == MC_MAJOR_MN;
    g.a_batched = (p->a_batch_stride != 0 && p->batch > 1); g.b_batched = (p->b_batch_stride != 0 && p->batch > 1);

    const int sms = sm_count();
    const bool linear_epi = p->act == MC_ACT_NONE && p->zout == nullptr && p->R == nullptr &&
                            p->bias_mode == MC_BIAS_NONE && p->c_dtype == MC_F32 && p->accumulate;
    const int heavy = p->act != MC_ACT_NONE;
    // tile width / split-K / cluster selection by a small cost model
    const int tiles_m_all = (int)ceil_div(p->M, BM);
    static const int allow_cluster = env_int("MC_GEMM_CLUSTER", 1);
    static const int allow_2cta_sel = env_int("MC_GEMM_2CTA", 1);
    int best_bn = 0, best_split = 1, best_cluster = 1;
    double best = 1e300;
    const int n_cap = (int)(ceil_div(p->N, 32) * 32);
    for (int cl = 1; cl <= (allow_cluster && tiles_m_all >= 2 && !dual ? 2 : 1); ++cl) {
        for (int bn = dual ? 128 : 256; bn >= 32; bn -= 32) {
            if (bn > n_cap && bn != 32) continue;
            if (cl == 2 && g.b_mn && bn % 128 != 0 && allow_2cta_sel) continue;   // pair mode needs whole 64-column groups per CTA
            int max_split = 1;
            if (p->split_k > 1) max_split = (int)p->split_k;
            else if (p->split_k == 0 && linear_epi) max_split = 64;
            // any split factor, not only powers of two: 16 tiles x 9 slices fill two rounds of 74 pair slots where
            // 16 x 4 leaves 10 of them idle and 16 x 8 needs two rounds of longer slices
            static const int any_split = env_int("MC_GEMM_ANY_SPLIT", 1);
            for (int sp = (p->split_k > 1 ? max_split : 1); sp <= max_split; sp = any_split ? sp + 1 : sp * 2) {
                if (sp > g.kb_total) break;
                double c = model_cost(p->M, p->N, g.out_batch, g.kb_total, bn, sp, sms, heavy, cl);
                if (sp > 1) c += 800.0;  // atomics
                if (c < best) { best = c; best_bn = bn; best_split = sp; best_cluster = cl; }
            }
        }
    }
    // tuning / debugging overrides (read on every call; used by tools/gemm_bench.py sweeps)
    {
        const int f_bn = env_int("MC_GEMM_BN", 0), f_sp = env_int("MC_GEMM_SPLIT", 0), f_cl = env_int("MC_GEMM_CL", 0);
        if (f_bn > 0 && f_bn <= (dual ? 128 : 256) && f_bn % 32 == 0) best_bn = f_bn;
        if (f_sp > 0 && (f_sp == 1 || linear_epi) && f_sp <= g.kb_total) best_split = f_sp;
        if (f_cl == 1 || (f_cl == 2 && tiles_m_all >= 2 && !dual)) best_cluster = f_cl;
    }
    MC_CHECK(best_bn > 0, "gemm: no tile configuration");
    if (best_split > 1) MC_CHECK(linear_epi || p->split_k > 1, "gemm: split-K needs a linear fp32 accumulate epilogue");
    g.BN = best_bn;
    g.split_k = best_split;
    g.cluster = best_cluster;
    g.tiles_m = tiles_m_all;
    g.tiles_m_eff = (int)ceil_div(g.tiles_m, g.cluster);
    g.tiles_n = (int)ceil_div(p->N, g.BN);
    const long long nt = (long long)g.tiles_m_eff * g.tiles_n * g.out_batch * g.split_k;   // work items per cluster
    MC_CHECK(nt < (1ll << 31), "gemm: too many tiles");
    g.num_tiles = (int)nt;
    static const int allow_2cta = env_int("MC_GEMM_2CTA", 1);
    g.two_cta = (allow_2cta && g.cluster == 2 && g.BN % 32 == 0 && (!g.b_mn || g.BN % 128 == 0)) ? 1 : 0;
    g.b_box_rows = g.BN / g.cluster;
    static const int prod2 = env_int("MC_GEMM_PROD2", 1);
    g.prod2 = prod2 ? 1 : 0;
    g.dbg_skip = env_int("MC_GEMM_DEBUG_SKIP", 0);
    {
        // Operands that stream from DRAM: the A operand of every GEMM (activations), and B too in the weight-gradient
        // GEMMs (both operands are activations).  Weights (the B operand of forward / dgrad, 1-5 MB) stay L2-resident.
        static const int l2pf = env_int("MC_GEMM_L2PF", 0);
        g.l2pf_a = l2pf > 0 ? l2pf : 0;
        g.l2pf_b = (l2pf > 0 && linear_epi) ? l2pf : 0;
        if (dual) g.l2pf_a = g.l2pf_b = 0;
    }
    static const int exp_flags = env_int("MC_GEMM_EXP", 3);   // both measured wins (profiles/r2_first_experiments.txt); 0 = round-1 paths
    g.exp_flags = exp_flags;
    // bytes of B landing in ONE CTA's stage: the whole tile (single / multicast) or its half (pair mode)
    const int bn_cta = g.two_cta ? g.BN / 2 : g.BN;
    g.b_tx_bytes = g.b_mn ? (int)(ceil_div(bn_cta, 64) * kGroupBytes) : bn_cta * BK * 2;
    g.pair_bytes = (int)(kABytes + ceil_div(g.b_tx_bytes, 1024) * 1024);
    g.stage_bytes = g.pair_bytes * (dual ? 2 : 1);
    g.dual = dual ? 1 : 0;
    g.a2_mn = p->a2_major == MC_MAJOR_MN; g.b2_mn = p->b2_major == MC_MAJOR_MN;
    g.a2_batched = (p->a2_batch_stride != 0 && p->batch > 1); g.b2_batched = (p->b2_batch_stride != 0 && p->batch > 1);
    g.bias2 = p->bias2;
    // epilogue kind: decided before the smem split because the TMA-store epilogues need a staging area
    const bool c_bf16 = p->c_dtype == MC_BF16;
    const bool atomic = best_split > 1;
    int epi = EPI_GENERIC;
    if (p->act == MC_ACT_GELU && c_bf16 && p->R == nullptr && !p->accumulate && !atomic) epi = EPI_ACT_FWD;
    else if (p->act == MC_ACT_GELU_BWD && c_bf16 && p->R == nullptr && p->zout == nullptr &&
             p->bias_mode == MC_BIAS_NONE && !p->accumulate && !atomic) epi = EPI_ACT_BWD;
    else if (p->act == MC_ACT_NONE && !c_bf16 && p->R != nullptr && p->zout == nullptr && !p->accumulate && !atomic)
        epi = EPI_RESID;
    else if (p->act == MC_ACT_NONE && !c_bf16 && p->R == nullptr && p->zout == nullptr && p->bias_mode == MC_BIAS_NONE)
        epi = EPI_PLAIN;
    if (dual) epi = EPI_ACT_BWD_DUAL;
    if (p->c_transposed) {
        MC_CHECK(p->act == MC_ACT_NONE && !c_bf16 && p->zout == nullptr && !p->accumulate && !atomic && p->row_remap == 0 &&
                     p->rowsum_out == nullptr,
                 "gemm: c_transposed supports plain fp32 stores with optional bias / residual only");
        epi = EPI_TRANS;
    }
    static const int force_generic = env_int("MC_GEMM_GENERIC_EPI", 0);
    static const int allow_tma_epi = env_int("MC_GEMM_TMA_EPI", 1);
    if (force_generic && epi != EPI_TRANS && epi != EPI_ACT_BWD_DUAL) epi = EPI_GENERIC;
    // TMA stores need 16-byte aligned bases / pitches and no row remapping
    auto al16 = [](const void* q, long long ld, long long bs, int esz) {
        return q == nullptr || ((reinterpret_cast<uintptr_t>(q) % 16 == 0) && (ld * esz) % 16 == 0 && (bs * esz) % 16 == 0);
    };
    g.tma_epi = allow_tma_epi && epi != EPI_GENERIC && epi != EPI_TRANS && p->row_remap == 0 && al16(p->R, p->ldr, p->r_batch_stride, 4) &&
                al16(p->C, p->ldc, p->c_batch_stride, c_bf16 ? 2 : 4) && al16(p->zout, p->ldz, p->z_batch_stride, 2) &&
                al16(p->zin, p->ldzin, p->zin_batch_stride, 2);
    static const int zdepth_env = env_int("MC_GEMM_ZDEPTH", 1);
    g.zdepth = (epi == EPI_ACT_BWD && zdepth_env >= 2) ? 2 : 1;
    // Two alternating OUTPUT staging buffers per epilogue warp (MC_GEMM_EPIBUF=2): the TMA store of chunk i reads its
    // tile while chunk i + 1 is computed (wait_group.read 1 instead of 0).  Costs 2-4 KB per warp = one ring stage;
    // the ring runs at 99 % tensor pipe with 3 stages (tools/ubench/ring_handover).
    static const int epibuf_env = env_int("MC_GEMM_EPIBUF", 1);
    g.epi_bufs = (epibuf_env >= 2 && epi != EPI_RESID && epi != EPI_GENERIC && epi != EPI_TRANS) ? 2 : 1;
    const int out_bytes = epi == EPI_PLAIN ? 4096 : (epi == EPI_ACT_FWD ? 4096 : 2048);   // bytes of one output buffer (ACT_FWD: C + Z)
    g.epi_warp_bytes = (int)kEpiWarpBytes + (g.zdepth - 1) * 2048;
    g.epi_alt_off = g.epi_warp_bytes;
    if (g.epi_bufs == 2) g.epi_warp_bytes += out_bytes;
    const int epi_bytes = g.tma_epi ? kEpiWarps * g.epi_warp_bytes : 0;
    const int smem_budget = 225 * 1024 - epi_bytes;
    g.stages = smem_budget / g.stage_bytes;
    if (g.stages > kMaxStages) g.stages = kMaxStages;
    MC_CHECK(g.stages >= 2, "gemm: not enough shared memory for 2 stages");
    g.epi_smem_off = g.stages * g.stage_bytes;
    // Always request >= 120 KB so that at most one CTA is resident per SM (each allocates all of TMEM).
    const size_t smem = (size_t)g.stages * g.stage_bytes + epi_bytes + 1024;

    g.C = p->C; g.c_bf16 = p->c_dtype == MC_BF16; g.ldc = p->ldc; g.c_bs = p->c_batch_stride;
    g.accumulate = p->accumulate; g.atomic = g.split_k > 1; g.row_remap = p->row_remap;
    g.bias = p->bias; g.bias_mode = p->bias_mode;
    g.zout = reinterpret_cast<__half*>(p->zout); g.ldz = p->ldz; g.z_bs = p->z_batch_stride;
    g.zin = reinterpret_cast<const __half*>(p->zin); g.ldzin = p->ldzin; g.zin_bs = p->zin_batch_stride;
    g.act = p->act; g.R = p->R; g.ldr = p->ldr; g.r_bs = p->r_batch_stride;
    g.rowsum_out = p->rowsum_out;
    g.rowstat_out = p->rowstat_out;
    MC_CHECK(p->rowstat_out == nullptr || (epi == EPI_RESID && g.tma_epi && g.out_batch == 1 && p->row_remap == 0 && !p->c_transposed),
             "gemm: rowstat_out needs the TMA residual epilogue (act NONE, fp32 C, aligned R), batch 1, no row_remap");
    g.c_transposed = p->c_transposed ? 1 : 0;
    MC_CHECK(p->rowsum_out == nullptr || epi == EPI_ACT_BWD || epi == EPI_ACT_BWD_DUAL || epi == EPI_GENERIC,
             "gemm: rowsum_out is supported with the GELU-backward and generic epilogues only");
    MC_CHECK(p->rowsum_out == nullptr || (p->row_remap == 0 && !g.k_spans_batch), "gemm: rowsum_out with row_remap / k_spans_batch");
    // vector (8-column) epilogue accesses need every touched row start to be 32-byte aligned
    auto al = [](const void* q, long long ld, long long bs, int esz) {
        return q == nullptr || ((reinterpret_cast<uintptr_t>(q) % 32 == 0) && (ld * esz) % 32 == 0 && (bs * esz) % 32 == 0);
    };
    g.vec_ok = al(p->C, p->ldc, p->c_batch_stride, g.c_bf16 ? 2 : 4) && al(p->zout, p->ldz, p->z_batch_stride, 2) &&
               al(p->zin, p->ldzin, p->zin_batch_stride, 2) && al(p->R, p->ldr, p->r_batch_stride, 4) &&
               (p->bias_mode != MC_BIAS_N || reinterpret_cast<uintptr_t>(p->bias) % 32 == 0) && (g.BN % 32 == 0);

    CUtensorMap tmA, tmB;
    int rc = make_operand_map(&tmA, p->A, p->a_major, p->M, p->K, p->lda, p->batch, p->a_batch_stride, BM, "A");
    if (rc != MC_OK) return rc;
    rc = make_operand_map(&tmB, p->B, p->b_major, p->N, p->K, p->ldb, p->batch, p->b_batch_stride, g.b_box_rows, "B");
    if (rc != MC_OK) return rc;

    CUtensorMap tmC, tmZ, tmA2, tmB2;
    memset(&tmC, 0, sizeof(tmC));
    memset(&tmZ, 0, sizeof(tmZ));
    memset(&tmA2, 0, sizeof(tmA2));
    memset(&tmB2, 0, sizeof(tmB2));
    if (dual) {
        MC_CHECK(g.tma_epi, "gemm: the recompute pair needs 16-byte aligned C rows (TMA-store epilogue)");
        rc = make_operand_map(&tmA2, p->A2, p->a2_major, p->M, p->K, p->lda2, p->batch, p->a2_batch_stride, BM, "A2");
        if (rc != MC_OK) return rc;
        rc = make_operand_map(&tmB2, p->B2, p->b2_major, p->N, p->K, p->ldb2, p->batch, p->b2_batch_stride, g.b_box_rows, "B2");
        if (rc != MC_OK) return rc;
    }
    if (g.tma_epi && epi == EPI_RESID) {   // only the residual goes through TMA (loaded); C is stored directly
        static const int rpromo = env_int("MC_GEMM_RPROMO", 0);
        rc = make_store_map(&tmZ, const_cast<float*>(p->R), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p->N, p->M, g.out_batch, p->ldr,
                            p->r_batch_stride, "R", rpromo);
        if (rc != MC_OK) return rc;
    } else if (g.tma_epi) {
        const int esz = c_bf16 ? 2 : 4;
        rc = make_store_map(&tmC, p->C, c_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, esz, p->N,
                            p->M, g.out_batch, p->ldc, p->c_batch_stride, "C");
        if (rc != MC_OK) return rc;
        if (p->zout != nullptr) {
            rc = make_store_map(&tmZ, p->zout, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p->N, p->M, g.out_batch, p->ldz,
                                p->z_batch_stride, "Z");
            if (rc != MC_OK) return rc;
        } else if (epi == EPI_ACT_BWD) {   // the saved pre-activations are LOADED through the same box geometry
            static const int zpromo = env_int("MC_GEMM_ZPROMO", 0);
            rc = make_store_map(&tmZ, const_cast<void*>(p->zin), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p->N, p->M, g.out_batch,
                                p->ldzin, p->zin_batch_stride, "Zin", zpromo);
            if (rc != MC_OK) return rc;
        }
    }
    // the direct (non-TMA) specialised paths use 256-bit accesses and need 32-byte aligned rows
    if (!g.tma_epi && epi != EPI_GENERIC && epi != EPI_TRANS && epi != EPI_ACT_BWD_DUAL && !(g.vec_ok && p->row_remap >= 0))
        epi = EPI_GENERIC;

    // Balanced persistent grid (MC_GEMM_BALANCED=1, off by default): the tiles are dealt in ceil(tiles / slots) rounds
    // whatever the grid, so the launch could take only as many slots as that round count needs (600 tiles on 74 pair
    // slots = 9 rounds = 67 slots) and leave the other SMs to the other streams.  Measured neutral on the two-stream
    // step (15.89 vs 15.94 ms) and slower together with the SM split (15.5 vs 15.0 ms), hence opt-in.
    static const int balanced = env_int("MC_GEMM_BALANCED", 0);
    int slots = sms / g.cluster;
    if (balanced && g.num_tiles > slots) {
        const int rounds = (g.num_tiles + slots - 1) / slots;
        slots = (g.num_tiles + rounds - 1) / rounds;
    }
    const int grid = (g.num_tiles < slots ? g.num_tiles : slots) * g.cluster;
    static const int debug = env_int("MC_GEMM_DEBUG", 0);
    if (debug)
        fprintf(stderr, "[gemm_tc] M=%d N=%d K=%d batch=%d kspan=%d a%db%d -> BN=%d split=%d cluster=%d 2cta=%d stages=%d epi=%d tma_epi=%d tiles=%d grid=%d\n",
                g.M, g.N, g.K, (int)p->batch, g.k_spans_batch, g.a_mn, g.b_mn, g.BN, g.split_k, g.cluster, g.two_cta, g.stages, epi,
                g.tma_epi, g.num_tiles, grid);
    const size_t smem_req = smem < 120 * 1024 ? 120 * 1024 : smem;
    switch (epi) {
        case EPI_ACT_FWD: return launch_gemm<EPI_ACT_FWD>(tmA, tmB, tmC, tmZ, tmA2, tmB2, g, grid, smem_req, stream);
        case EPI_RESID: return launch_gemm<EPI_RESID>(tmA, tmB, tmC, tmZ, tmA2, tmB2, g, grid, smem_req, stream);
        case EPI_ACT_BWD: return launch_gemm<EPI_ACT_BWD>(tmA, tmB, tmC, tmZ, tmA2, tmB2, g, grid, smem_req, stream);
        case EPI_PLAIN: return launch_gemm<EPI_PLAIN>(tmA, tmB, tmC, tmZ, tmA2, tmB2, g, grid, smem_req, stream);
        case EPI_TRANS: return launch_gemm<EPI_TRANS>(tmA, tmB, tmC, tmZ, tmA2, tmB2, g, grid, smem_req, stream);
        case EPI_ACT_BWD_DUAL: return launch_gemm_t<EPI_ACT_BWD_DUAL, false>(tmA, tmB, tmC, tmZ, tmA2, tmB2, g, grid, smem_req, stream);
        default: return launch_gemm<EPI_GENERIC>(tmA, tmB, tmC, tmZ, tmA2, tmB2, g, grid, smem_req, stream);
    }
}
