// Fused token-mixing MLP of a Mixer block (sm_100a): tcgen05 / TMEM / TMA, hidden activation never leaves the SM.
//
// Reference: MixerBlock.token_mix + token_mix_seq, training/clip/model.py:206-208,216,220-222
//     x = x + ( lin2( QuickGELU( lin1( LN1(x).permute(0,2,1) ) ) ) ).permute(0,2,1)
// and the backward autograd derives from it (training/training.py:170).  Per sample b, with U = LN1(X_b) [P x D]:
//     forward   Y  = X + W2 g(W1 U + b1) + b2
//     dgrad     dU = W1^T ( (W2^T dY) * g'(W1 U + b1) )
//     wgrad     dW2 += dY H1^T,  dW1 += dZ1 U^T,  db1 += rowsum(dZ1)          (H1, dZ1 recomputed per tile)
//
// The reference materialises the transposed activation and the [B, D, 4P] hidden tensor (three HBM round trips);
// here one persistent CTA per SM walks over (sample, 128-channel slab) tiles with D as the MMA M dimension:
//
//     Z^T [128 d x 4P]  = U_b^T  [128 d x P]  . W1^T [P x 4P]     A: MN-major view of the [P x D] activation tile
//                                                                   (the transpose lives in the UMMA descriptor)
//     H^T               = g(Z^T + b1)   TMEM -> registers -> bf16 -> 128B-swizzled smem (K-major A operand)
//     Y^T [128 d x P]   = H^T [128 d x 4P] . W2^T [4P x P]
//     y[b, p, d]        = x[b, p, d] + Y^T[d, p] + b2[p]           lane = d: every global access is a coalesced line
//
// Both weights stay resident in shared memory for the whole kernel as ONE [P x 4P] (row p, j contiguous) swizzled
// tile each (w1t[p][j] = W1[j][p], w2t[p][j] = W2[p][j]); the same bytes serve as the MN-major B operand of the
// "up" GEMMs (N = j, K = p) and as the K-major B operand of the "down" GEMMs (N = p, K = j).
//
//   warp 0      TMA producer: activation tiles (bf16 [P x 128] as two 64-wide swizzled groups) + L2 prefetch of the
//               fp32 residual tile
//   warp 1      MMA issuer (one thread)
//   warp 2      TMEM allocator
//   warp 3      optional TMA store of the bf16 H^T / dZ1^T tile ("spill", layout [B, D, 4P]) for an unfused consumer
//   warps 4-11  E1: TMEM -> bias + QuickGELU (or QuickGELU' product) -> bf16 -> smem operand tile
//   warps 12-19 E2: TMEM -> + residual + bias -> global (forward) / plain store (dgrad)
//   WGRAD mode: warps 12-19 idle; the weight-gradient accumulators stay in TMEM across all tiles of the CTA.
#include <string.h>

#include "common.cuh"

namespace mc {

namespace {

constexpr int kE1Warps = 8;
constexpr int kE2Warps = 8;
constexpr int kTmThreads = 128 + 32 * (kE1Warps + kE2Warps);   // 640
constexpr int kMaxAtoms = 5;                                   // hidden width 4P <= 320
constexpr int kMaxTmStages = 2;
constexpr uint32_t kAtomBytes = 128 * 128;                     // [128 rows x 64 bf16] swizzled K-major atom

enum { TM_FWD = 0, TM_DGRAD = 1, TM_WGRAD = 2 };

struct TmArgs {
    int B, P, D, H;
    int Ppad, Hpad, natoms;
    int tiles_d, num_tiles;
    int nseg, seg_c0[4], seg_w[4];
    int zcols;                // TMEM columns of one working buffer (multiple of 32)
    int stages;
    uint32_t grp_bytes;       // Ppad * 128: one [Ppad x 64] swizzled group of an activation / weight tile
    uint32_t a_grp_bytes;     // group pitch of the activation tiles in smem (>= grp_bytes; WGRAD: 16 KB, rows up to 128)
    uint32_t off_w1t, off_w2t, off_h, off_h2, off_stage, stage_bytes, off_b1, off_b2;
    const __nv_bfloat16* w1;
    int ld1;
    const __nv_bfloat16* w2;
    int ld2;
    const float* b1;
    const float* b2;
    const float* x;
    float* y;
    int spill;
    // WGRAD
    int j0, jw;               // hidden slice [j0, j0 + jw) of this launch's CTAs comes from blockIdx (see kernel)
    int nslices, slice_w;
    float* gw1;               // [4P, ldg1] fp32, +=
    int ldg1;
    float* gw2;               // [P, ldg2] fp32, +=
    int ldg2;
    float* gb1;               // [4P] fp32, +=
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tm_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tm_tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tm_tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tm_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tm_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tm_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// QuickGELU (model.py:175-177) on packed pairs: x * sigmoid(1.702 x) = hx + hx * tanh(0.851 x)
__device__ __forceinline__ float2 tm_gelu2(float2 x) {
    const float2 a = __fmul2_rn(x, make_float2(0.851f, 0.851f));
    const float2 t = make_float2(tanh_approx(a.x), tanh_approx(a.y));
    const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    return __ffma2_rn(hx, t, hx);
}
// sigmoid from ex2 + rcp (~1 ulp each): the derivative multiplies every gradient that flows through the block
__device__ __forceinline__ float2 tm_sigmoid2(float2 z) {
    const float2 a = __fmul2_rn(z, make_float2(-2.4554669595930157f, -2.4554669595930157f));   // -1.702 * log2(e) * z
    float ex, ey, sx, sy;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(a.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(a.y));
    const float2 den = __fadd2_rn(make_float2(ex, ey), make_float2(1.0f, 1.0f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sx) : "f"(den.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sy) : "f"(den.y));
    return make_float2(sx, sy);
}
__device__ __forceinline__ float2 tm_gelu_grad_from_s(float2 z, float2 s) {
    const float2 w = __fmul2_rn(z, make_float2(kGeluA, kGeluA));
    const float2 nw = __fmul2_rn(z, make_float2(-kGeluA, -kGeluA));
    const float2 t1 = __ffma2_rn(nw, s, w);      // w (1 - s)
    return __ffma2_rn(t1, s, s);                 // s + w s (1 - s)
}

// Weights -> resident swizzled tiles.  Element (p, j) of a tile lives at
//   (j / 64) * grp_bytes + p * 128 + (((j % 64) / 8) ^ (p % 8)) * 16 + (j % 8) * 2
// which is exactly what TMA's SWIZZLE_128B would produce for a [Ppad x 64] box of a [P x 4P] row-major matrix.
// Only the hidden slice [jbase, jbase + natoms * 64) is loaded (tile-local column jl = j - jbase).
__device__ __forceinline__ void load_weight_tiles(const TmArgs& g, uint32_t w1t, uint32_t w2t, int jbase, int natoms) {
    const int chunks_per_row = natoms * 8;
    const int total = g.Ppad * chunks_per_row;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int p = idx / chunks_per_row, cj = idx - p * chunks_per_row;
        const int grp = cj >> 3, c = cj & 7, jl = cj * 8, j = jbase + jl;
        uint32_t v1[4] = {0u, 0u, 0u, 0u}, v2[4] = {0u, 0u, 0u, 0u};
        if (p < g.P) {
            unsigned short t1[8], t2[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const bool ok = j + e < g.H;
                t1[e] = ok ? reinterpret_cast<const unsigned short*>(g.w1)[(long long)(j + e) * g.ld1 + p] : (unsigned short)0;
                t2[e] = ok ? reinterpret_cast<const unsigned short*>(g.w2)[(long long)p * g.ld2 + j + e] : (unsigned short)0;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                v1[e] = (uint32_t)t1[2 * e] | ((uint32_t)t1[2 * e + 1] << 16);
                v2[e] = (uint32_t)t2[2 * e] | ((uint32_t)t2[2 * e + 1] << 16);
            }
        }
        const uint32_t off = (uint32_t)grp * g.grp_bytes + (uint32_t)p * 128u + (uint32_t)((c ^ (p & 7)) << 4);
        tm_sts128(w1t + off, v1[0], v1[1], v1[2], v1[3]);
        tm_sts128(w2t + off, v2[0], v2[1], v2[2], v2[3]);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kTmThreads, 1)
token_mix_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmDY,
                 const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmS, const TmArgs g) {
    extern __shared__ uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t u_full[kMaxTmStages];
    __shared__ __align__(8) uint64_t u_empty[kMaxTmStages];
    __shared__ __align__(8) uint64_t z_full, z_empty, h_empty;
    __shared__ __align__(8) uint64_t h_full[kMaxAtoms];
    __shared__ __align__(8) uint64_t y_full[2];
    __shared__ __align__(8) uint64_t y_empty[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(dyn_smem) + 1023u) & ~1023u;
    const uint32_t w1t = base + g.off_w1t, w2t = base + g.off_w2t, hbuf = base + g.off_h;
    float* b1s = reinterpret_cast<float*>(dyn_smem + (base - smem_u32(dyn_smem)) + g.off_b1);
    float* b2s = reinterpret_cast<float*>(dyn_smem + (base - smem_u32(dyn_smem)) + g.off_b2);
    constexpr int kYBufs = MODE == TM_FWD ? 2 : 1;
    // WGRAD: CTAs are dealt round-robin to the hidden slices; every slice walks over all tiles
    const int slice = MODE == TM_WGRAD ? (int)(blockIdx.x % g.nslices) : 0;
    const int jbase = MODE == TM_WGRAD ? slice * g.slice_w : 0;
    const int work0 = MODE == TM_WGRAD ? (int)(blockIdx.x / g.nslices) : (int)blockIdx.x;
    const int work_stride = MODE == TM_WGRAD ? (int)(gridDim.x / g.nslices) : (int)gridDim.x;
    const bool active = MODE != TM_WGRAD || (int)blockIdx.x < work_stride * g.nslices;
    // hidden columns handled by this CTA (tile-local: column c <-> hidden unit jbase + c)
    int my_hpad = g.Hpad, my_atoms = g.natoms;
    if (MODE == TM_WGRAD) {
        my_hpad = g.Hpad - jbase < g.slice_w ? g.Hpad - jbase : g.slice_w;
        my_atoms = (my_hpad + 63) / 64;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmU);
        if (MODE != TM_FWD) tma_prefetch_desc(&tmDY);
        if (MODE == TM_FWD) tma_prefetch_desc(&tmX);
        if (g.spill) tma_prefetch_desc(&tmS);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kMaxTmStages; ++s) {
            mbar_init(smem_u32(&u_full[s]), 1);
            mbar_init(smem_u32(&u_empty[s]), 1);
        }
        mbar_init(smem_u32(&z_full), 1);
        mbar_init(smem_u32(&z_empty), kE1Warps);
        mbar_init(smem_u32(&h_empty), g.spill ? 2 : 1);
        for (int a = 0; a < kMaxAtoms; ++a) mbar_init(smem_u32(&h_full[a]), kE1Warps);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&y_full[s]), 1);
            mbar_init(smem_u32(&y_empty[s]), kE2Warps);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(&tmem_base_smem), 512);
        tmem_relinquish();
    }
    // resident operands
    load_weight_tiles(g, w1t, w2t, jbase, MODE == TM_WGRAD ? my_atoms : g.natoms);
    for (int i = threadIdx.x; i < g.natoms * 64; i += blockDim.x) b1s[i] = (jbase + i < g.H) ? g.b1[jbase + i] : 0.f;
    if (MODE == TM_FWD)
        for (int i = threadIdx.x; i < g.Ppad; i += blockDim.x) b2s[i] = i < g.P ? g.b2[i] : 0.f;
    if (MODE == TM_WGRAD) {
        // rows Ppad .. 127 of every U group are constant: row Ppad is all ones (its accumulator lane collects
        // db1 = sum_d dZ1), the others zero.  TMA only ever rewrites rows < Ppad of a group.
        const uint32_t stage0 = base + g.off_stage;
        const int rows_c = 128 - g.Ppad;
        for (int s = 0; s < g.stages; ++s)
            for (int grp = 0; grp < 2; ++grp)
                for (int i = threadIdx.x; i < rows_c * 8; i += blockDim.x) {
                    const int r = g.Ppad + (i >> 3);
                    const uint32_t v = r == g.Ppad ? 0x3f803f80u : 0u;     // bf16 1.0 pairs
                    tm_sts128(stage0 + s * g.stage_bytes + grp * g.a_grp_bytes + r * 128 + ((i & 7) << 4), v, v, v, v);
                }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    // TMEM column map
    const uint32_t col_z = 0, col_dh = g.zcols;
    const uint32_t col_y = MODE == TM_FWD ? (uint32_t)g.zcols : 2u * g.zcols;          // FWD: Y0, Y1; DGRAD: dU
    const uint32_t col_acc2 = 2u * g.zcols, col_acc1 = 2u * g.zcols + g.slice_w;       // WGRAD: dW2 / dW1^T accumulators

    if (!active) {
        // nothing to do for this CTA (WGRAD with a grid that is not a multiple of the slice count)
    } else if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (int t = work0; t < g.num_tiles; t += work_stride) {
                const int b = t / g.tiles_d, d0 = (t - b * g.tiles_d) * 128;
                mbar_wait(smem_u32(&u_empty[st]), ph ^ 1u);
                const uint32_t bar = smem_u32(&u_full[st]);
                const uint32_t dst = base + g.off_stage + st * g.stage_bytes;
                mbar_arrive_expect_tx(bar, (MODE == TM_FWD ? 2u : 4u) * g.grp_bytes);
                tma_load_3d(dst, &tmU, bar, d0, 0, b);
                tma_load_3d(dst + g.a_grp_bytes, &tmU, bar, d0 + 64, 0, b);
                if (MODE != TM_FWD) {
                    tma_load_3d(dst + 2 * g.a_grp_bytes, &tmDY, bar, d0, 0, b);
                    tma_load_3d(dst + 2 * g.a_grp_bytes + g.grp_bytes, &tmDY, bar, d0 + 64, 0, b);
                }
                if (MODE == TM_FWD) {
                    // the fp32 residual tile is read by the E2 warps straight from global memory: pull it into L2 now
#pragma unroll
                    for (int i = 0; i < 4; ++i) tm_tma_prefetch_3d(&tmX, d0 + 32 * i, 0, b);
                }
                if (++st == (uint32_t)g.stages) {
                    st = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t st = 0, ph = 0, zph = 0, hph = 0, yb = 0, yph[2] = {0u, 0u};
            const uint32_t idesc_dn = make_idesc_bf16(128, g.Ppad, 0, 0);
            bool first_tile = true;
            for (int t = work0; t < g.num_tiles; t += work_stride) {
                mbar_wait(smem_u32(&u_full[st]), ph);
                tc_fence_after();
                const uint32_t u_base = base + g.off_stage + st * g.stage_bytes;
                const uint32_t dy_base = u_base + 2 * g.a_grp_bytes;
                // ---- "up" GEMMs: Z^T = U^T W1^T (and dH^T = dY^T W2), one hidden segment at a time ----
                for (int s = 0; s < g.nseg; ++s) {
                    mbar_wait(smem_u32(&z_empty), zph ^ 1u);
                    zph ^= 1u;
                    tc_fence_after();
                    int c = 0;
                    const int sw = MODE == TM_WGRAD ? my_hpad : g.seg_w[s];
                    while (c < sw) {
                        int w = sw - c;
                        if (w > 256) w = 192;
                        const uint32_t idesc_up = make_idesc_bf16(128, w, 1, 1);
                        const uint32_t grp0 = (uint32_t)((g.seg_c0[s] + c) >> 6);
                        for (int ks = 0; ks < g.Ppad / 16; ++ks) {
                            const uint64_t ad = make_sdesc_sw128(u_base + ks * 2048, g.a_grp_bytes, 1024u);
                            const uint64_t bd = make_sdesc_sw128(w1t + grp0 * g.grp_bytes + ks * 2048, g.grp_bytes, 1024u);
                            umma_ss(tmem_base + col_z + c, ad, bd, idesc_up, ks > 0 ? 1u : 0u);
                            if (MODE != TM_FWD) {
                                const uint64_t ad2 = make_sdesc_sw128(dy_base + ks * 2048, g.grp_bytes, 1024u);
                                const uint64_t bd2 = make_sdesc_sw128(w2t + grp0 * g.grp_bytes + ks * 2048, g.grp_bytes, 1024u);
                                umma_ss(tmem_base + col_dh + c, ad2, bd2, idesc_up, ks > 0 ? 1u : 0u);
                            }
                        }
                        c += w;
                    }
                    if (MODE != TM_WGRAD && s == g.nseg - 1) umma_commit(smem_u32(&u_empty[st]));
                    umma_commit(smem_u32(&z_full));
                }
                // ---- "down" GEMMs ----
                if (MODE == TM_WGRAD) {
                    // dW2[p, j] += sum_d dY[p, d] H^T[d, j];  dW1^T[p, j] += sum_d U[p, d] dZ1^T[d, j]   (K = 128 channels)
                    const uint32_t idesc_w = make_idesc_bf16(128, my_hpad, 0, 1);
                    const uint32_t h2buf = base + g.off_h2;
                    for (int a = 0; a < my_atoms; ++a) {
                        mbar_wait(smem_u32(&h_full[a]), hph);
                    }
                    tc_fence_after();
                    for (int ks = 0; ks < 8; ++ks) {
                        // A (K-major): rows p, 64-channel atoms = the activation groups; k-step 16 channels = 32 B
                        const uint32_t aoff = (uint32_t)(ks >> 2) * g.a_grp_bytes + (uint32_t)(ks & 3) * 32u;
                        const uint64_t a_dy = make_sdesc_sw128(dy_base + (uint32_t)(ks >> 2) * g.grp_bytes + (uint32_t)(ks & 3) * 32u, 16u, 1024u);
                        const uint64_t a_u = make_sdesc_sw128(u_base + aoff, 16u, 1024u);
                        // B (MN-major): K = channel rows of the [128 d x 64 j] atoms, 16 rows = 2048 B; groups of 64 j = atoms
                        const uint64_t b_h = make_sdesc_sw128(hbuf + ks * 2048, kAtomBytes, 1024u);
                        const uint64_t b_dz = make_sdesc_sw128(h2buf + ks * 2048, kAtomBytes, 1024u);
                        umma_ss(tmem_base + col_acc2, a_dy, b_h, idesc_w, (!first_tile || ks > 0) ? 1u : 0u);
                        umma_ss(tmem_base + col_acc1, a_u, b_dz, idesc_w, (!first_tile || ks > 0) ? 1u : 0u);
                    }
                    hph ^= 1u;
                    umma_commit(smem_u32(&u_empty[st]));
                    umma_commit(smem_u32(&h_empty));
                } else {
                    mbar_wait(smem_u32(&y_empty[yb]), yph[yb] ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + col_y + yb * g.Ppad;
                    const uint32_t wdn = MODE == TM_FWD ? w2t : w1t;
                    for (int a = 0; a < g.natoms; ++a) {
                        mbar_wait(smem_u32(&h_full[a]), hph);
                        tc_fence_after();
                        const int ksteps = (g.Hpad - a * 64) >= 64 ? 4 : (g.Hpad - a * 64) / 16;
                        for (int kk = 0; kk < ksteps; ++kk) {
                            const uint64_t ad = make_sdesc_sw128(hbuf + a * kAtomBytes + kk * 32, 16u, 1024u);
                            const uint64_t bd = make_sdesc_sw128(wdn + a * g.grp_bytes + kk * 32, 16u, 1024u);
                            umma_ss(d_tmem, ad, bd, idesc_dn, (a > 0 || kk > 0) ? 1u : 0u);
                        }
                    }
                    hph ^= 1u;
                    umma_commit(smem_u32(&h_empty));
                    umma_commit(smem_u32(&y_full[yb]));
                    yph[yb] ^= 1u;
                    yb = (yb + 1) % kYBufs;
                }
                first_tile = false;
                if (++st == (uint32_t)g.stages) {
                    st = 0;
                    ph ^= 1u;
                }
            }
            if (MODE == TM_WGRAD) umma_commit(smem_u32(&y_full[0]));   // accumulators final
        }
    } else if (warp == 3) {
        // ===================== spill: TMA store of the bf16 operand tile =====================
        if (g.spill && lane == 0 && MODE != TM_WGRAD) {
            uint32_t hph = 0;
            for (int t = work0; t < g.num_tiles; t += work_stride) {
                const int b = t / g.tiles_d, d0 = (t - b * g.tiles_d) * 128;
                for (int a = 0; a < g.natoms; ++a) {
                    mbar_wait(smem_u32(&h_full[a]), hph);
                    tm_tma_store_3d(&tmS, hbuf + a * kAtomBytes, a * 64, d0, b);
                }
                tm_bulk_commit();
                tm_bulk_wait_read0();
                mbar_arrive(smem_u32(&h_empty));
                hph ^= 1u;
            }
            tm_bulk_wait_all();
        }
    } else if (warp >= 4 && warp < 4 + kE1Warps) {
        // ===================== E1: hidden activation TMEM -> bf16 smem operand =====================
        const int e = warp - 4;
        const int q = e & 3, par = e >> 2;
        const int r = q * 32 + lane;                         // tile row = channel d0 + r
        const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        uint32_t zfph = 0, heph = 0;
        for (int t = work0; t < g.num_tiles; t += work_stride) {
            for (int s = 0; s < g.nseg; ++s) {
                mbar_wait(smem_u32(&z_full), zfph);
                zfph ^= 1u;
                tc_fence_after();
                if (s == 0) {
                    mbar_wait(smem_u32(&h_empty), heph ^ 1u);
                    heph ^= 1u;
                }
                const int c0 = g.seg_c0[s];
                const int sw_cols = MODE == TM_WGRAD ? my_hpad : g.seg_w[s];
                const int a_begin = c0 >> 6, a_end = (c0 + sw_cols + 63) >> 6;
                for (int a = a_begin; a < a_end; ++a) {
                    const int col = a * 64 + par * 32;       // tile-local hidden column of this warp's chunk
                    const int rel = col - c0;
                    if (rel < sw_cols) {
                        const uint32_t dst = hbuf + a * kAtomBytes + row_off;
                        if (MODE == TM_FWD) {
                            uint32_t v[32];
                            tmem_ld32(t_lane + col_z + rel, v);
                            tmem_ld_wait();
                            uint32_t o[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float2 bb = *reinterpret_cast<const float2*>(b1s + col + 2 * i);
                                const float2 z = __fadd2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), bb);
                                const float2 h = tm_gelu2(z);
                                o[i] = pack_bf16x2(h.x, h.y);
                            }
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj)
                                tm_sts128(dst + (((uint32_t)(par * 4 + jj) ^ sw) << 4), o[4 * jj], o[4 * jj + 1], o[4 * jj + 2], o[4 * jj + 3]);
                        } else {
                            const uint32_t dst2 = base + g.off_h2 + a * kAtomBytes + row_off;
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {
                                uint32_t zv[16], dv[16];
                                tmem_ld16(t_lane + col_z + rel + hf * 16, zv);
                                tmem_ld16(t_lane + col_dh + rel + hf * 16, dv);
                                tmem_ld_wait();
                                uint32_t o[8], oh[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float2 bb = *reinterpret_cast<const float2*>(b1s + col + hf * 16 + 2 * i);
                                    const float2 z = __fadd2_rn(make_float2(__uint_as_float(zv[2 * i]), __uint_as_float(zv[2 * i + 1])), bb);
                                    const float2 sg = tm_sigmoid2(z);
                                    const float2 gp = tm_gelu_grad_from_s(z, sg);
                                    const float2 dz = __fmul2_rn(make_float2(__uint_as_float(dv[2 * i]), __uint_as_float(dv[2 * i + 1])), gp);
                                    o[i] = pack_bf16x2(dz.x, dz.y);
                                    if (MODE == TM_WGRAD) {
                                        const float2 h = __fmul2_rn(z, sg);
                                        oh[i] = pack_bf16x2(h.x, h.y);
                                    }
                                }
                                const uint32_t cc = (uint32_t)(par * 4 + hf * 2);
                                if (MODE == TM_WGRAD) {
                                    // hbuf holds H^T, h2buf holds dZ1^T
                                    tm_sts128(dst + ((cc ^ sw) << 4), oh[0], oh[1], oh[2], oh[3]);
                                    tm_sts128(dst + (((cc + 1) ^ sw) << 4), oh[4], oh[5], oh[6], oh[7]);
                                    tm_sts128(dst2 + ((cc ^ sw) << 4), o[0], o[1], o[2], o[3]);
                                    tm_sts128(dst2 + (((cc + 1) ^ sw) << 4), o[4], o[5], o[6], o[7]);
                                } else {
                                    tm_sts128(dst + ((cc ^ sw) << 4), o[0], o[1], o[2], o[3]);
                                    tm_sts128(dst + (((cc + 1) ^ sw) << 4), o[4], o[5], o[6], o[7]);
                                }
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&h_full[a]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&z_empty));
            }
        }
        if (MODE == TM_WGRAD) {
            // ---- final: accumulators -> global gradients.  Lane = token p (or the ones-row Ppad -> db1) ----
            mbar_wait(smem_u32(&y_full[0]), 0u);
            tc_fence_after();
            const int p = r;
            for (int c = par * 32; c < my_hpad; c += 64) {
                uint32_t v2[32], v1[32];
                tmem_ld32(t_lane + col_acc2 + c, v2);
                tmem_ld32(t_lane + col_acc1 + c, v1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int j = jbase + c + i;
                    if (j < g.H) {
                        if (p < g.P) {
                            atomicAdd(g.gw2 + (long long)p * g.ldg2 + j, __uint_as_float(v2[i]));
                            atomicAdd(g.gw1 + (long long)j * g.ldg1 + p, __uint_as_float(v1[i]));
                        } else if (p == g.Ppad) {
                            atomicAdd(g.gb1 + j, __uint_as_float(v1[i]));
                        }
                    }
                }
            }
        }
    } else if (warp >= 4 + kE1Warps) {
        // ===================== E2: output accumulator -> global =====================
        if (MODE != TM_WGRAD) {
            const int e2 = warp - 4 - kE1Warps;
            const int q = e2 & 3, hsel = e2 >> 2;
            const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
            const int nch = g.Ppad / 16;
            uint32_t yb = 0, yph[2] = {0u, 0u};
            for (int t = work0; t < g.num_tiles; t += work_stride) {
                const int b = t / g.tiles_d, d0 = (t - b * g.tiles_d) * 128;
                const long long gbase = (long long)b * g.P * g.D + d0 + q * 32 + lane;
                float xr[3][16];
                if (MODE == TM_FWD) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int ch = hsel + 2 * k;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int p = ch * 16 + i;
                            xr[k][i] = (ch < nch && p < g.P) ? __ldg(g.x + gbase + (long long)p * g.D) : 0.f;
                        }
                    }
                }
                mbar_wait(smem_u32(&y_full[yb]), yph[yb]);
                yph[yb] ^= 1u;
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int ch = hsel + 2 * k;
                    if (ch < nch) {
                        uint32_t v[16];
                        tmem_ld16(t_lane + col_y + yb * g.Ppad + ch * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int p = ch * 16 + i;
                            if (p < g.P) {
                                float o = __uint_as_float(v[i]);
                                if (MODE == TM_FWD) o += b2s[p] + xr[k][i];
                                g.y[gbase + (long long)p * g.D] = o;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&y_empty[yb]));
                yb = (yb + 1) % kYBufs;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---- host ----------------------------------------------------------------------------------------------
typedef CUresult (*TmEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

TmEncodeFn tm_encode_fn() {
    static TmEncodeFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TmEncodeFn>(p);
    }
    return fn;
}

// 3-D map over a dense [B][rows][cols] tensor (cols contiguous, row pitch ld elements)
int tm_make_map(CUtensorMap* map, const void* ptr, CUtensorMapDataType dt, int esz, int64_t cols, int64_t rows, int64_t batch,
                int64_t ld, int64_t batch_stride, int box_cols, int box_rows, CUtensorMapSwizzle swz, const char* name) {
    TmEncodeFn enc = tm_encode_fn();
    MC_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    MC_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "token_mix %s: base pointer must be 16-byte aligned", name);
    MC_CHECK((ld * esz) % 16 == 0 && (batch_stride * esz) % 16 == 0, "token_mix %s: pitches must be multiples of 16 bytes", name);
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t gstride[2] = {(cuuint64_t)ld * esz, (cuuint64_t)batch_stride * esz};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1}, estr[3] = {1, 1, 1};
    CUresult r = enc(map, dt, 3, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MC_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(token_mix %s) failed with %d", name, (int)r);
    return MC_OK;
}

template <int MODE>
int tm_launch(const CUtensorMap& tmU, const CUtensorMap& tmDY, const CUtensorMap& tmX, const CUtensorMap& tmS,
              const TmArgs& g, int grid, size_t smem, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        MC_CUDA(cudaFuncSetAttribute(token_mix_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        attr_set = true;
    }
    token_mix_kernel<MODE><<<grid, kTmThreads, smem, stream>>>(tmU, tmDY, tmX, tmS, g);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

int round_up_i(int x, int m) { return (x + m - 1) / m * m; }

int tm_run(const mc_token_mix_params* p, int mode, cudaStream_t stream) {
    MC_CHECK(p != nullptr, "token_mix: null params");
    MC_CHECK(p->B > 0 && p->P > 0 && p->D > 0, "token_mix: empty problem");
    MC_CHECK(mc_token_mix_supported(p->P, p->D), "token_mix: unsupported shape P=%lld D=%lld (need P <= 80, D %% 128 == 0)",
             (long long)p->P, (long long)p->D);
    MC_CHECK(p->u && p->w1 && p->w2 && p->b1, "token_mix: null operand");
    TmArgs g{};
    g.B = (int)p->B; g.P = (int)p->P; g.D = (int)p->D; g.H = 4 * g.P;
    g.Ppad = round_up_i(g.P, 16);
    g.Hpad = round_up_i(g.H, 16);
    g.natoms = (g.Hpad + 63) / 64;
    g.tiles_d = g.D / 128;
    MC_CHECK((long long)g.B * g.tiles_d < (1ll << 31), "token_mix: too many tiles");
    g.num_tiles = g.B * g.tiles_d;
    g.grp_bytes = (uint32_t)g.Ppad * 128u;
    g.a_grp_bytes = mode == TM_WGRAD ? kAtomBytes : g.grp_bytes;
    g.w1 = reinterpret_cast<const __nv_bfloat16*>(p->w1); g.ld1 = (int)p->ld1;
    g.w2 = reinterpret_cast<const __nv_bfloat16*>(p->w2); g.ld2 = (int)p->ld2;
    g.b1 = p->b1; g.b2 = p->b2; g.x = p->x; g.y = p->y;
    g.spill = (p->spill != nullptr && mode != TM_WGRAD) ? 1 : 0;
    MC_CHECK(g.ld1 >= g.P && g.ld2 >= g.H, "token_mix: weight pitches too small");

    const int sms = sm_count();
    int grid = g.num_tiles < sms ? g.num_tiles : sms;
    int natoms_smem = g.natoms;
    if (mode == TM_WGRAD) {
        // accumulators 2 x slice_w + working buffers 2 x slice_w TMEM columns: slice_w = 128
        g.slice_w = 128;
        g.nslices = (g.Hpad + g.slice_w - 1) / g.slice_w;
        g.nseg = 1; g.seg_c0[0] = 0; g.seg_w[0] = g.slice_w;
        g.zcols = g.slice_w;
        natoms_smem = g.slice_w / 64;
        MC_CHECK(p->gw1 && p->gw2 && p->gb1 && p->dy, "token_mix wgrad: null operand");
        g.gw1 = p->gw1; g.ldg1 = (int)p->ldg1; g.gw2 = p->gw2; g.ldg2 = (int)p->ldg2; g.gb1 = p->gb1;
        grid = sms / g.nslices * g.nslices;
        if (grid > g.num_tiles * g.nslices) grid = g.num_tiles * g.nslices;
    } else {
        // hidden segments: the working TMEM buffers (Z, and dH in dgrad) share 512 columns with the output tile(s)
        const int cap = mode == TM_FWD ? 512 - 2 * g.Ppad : (512 - g.Ppad) / 2;
        if (round_up_i(g.Hpad, 32) <= cap) {
            g.nseg = 1; g.seg_c0[0] = 0; g.seg_w[0] = g.Hpad;
        } else {
            const int wseg = cap / 64 * 64;
            g.nseg = 0;
            for (int c = 0; c < g.Hpad; c += wseg) {
                MC_CHECK(g.nseg < 4, "token_mix: too many hidden segments");
                g.seg_c0[g.nseg] = c;
                g.seg_w[g.nseg] = g.Hpad - c < wseg ? g.Hpad - c : wseg;
                ++g.nseg;
            }
        }
        g.zcols = 0;
        for (int s = 0; s < g.nseg; ++s) g.zcols = g.zcols > round_up_i(g.seg_w[s], 32) ? g.zcols : round_up_i(g.seg_w[s], 32);
    }
    // shared memory plan (all offsets multiples of 1024)
    uint32_t off = 0;
    g.off_w1t = off; off += (uint32_t)natoms_smem * g.grp_bytes;
    off = (off + 1023u) & ~1023u;
    g.off_w2t = off; off += (uint32_t)natoms_smem * g.grp_bytes;
    off = (off + 1023u) & ~1023u;
    g.off_h = off; off += (uint32_t)natoms_smem * kAtomBytes;
    g.off_h2 = off;
    if (mode == TM_WGRAD) off += (uint32_t)natoms_smem * kAtomBytes;
    g.off_b1 = off; off += (uint32_t)g.natoms * 64u * 4u;
    g.off_b2 = off; off += (uint32_t)round_up_i(g.Ppad * 4, 1024);
    off = (off + 1023u) & ~1023u;
    g.off_stage = off;
    g.stage_bytes = mode == TM_FWD ? 2u * g.grp_bytes : 2u * g.a_grp_bytes + 2u * g.grp_bytes;
    g.stage_bytes = (g.stage_bytes + 1023u) & ~1023u;
    // WGRAD reads the Ppad-row dY groups as a 128-row K-major operand: the rows past the last stage feed ignored
    // accumulator lanes but must lie inside the allocation
    const uint32_t slack = mode == TM_WGRAD ? (uint32_t)(128 - g.Ppad) * 128u : 0u;
    const uint32_t budget = 226u * 1024u - 1024u - slack;
    MC_CHECK(off + g.stage_bytes <= budget, "token_mix: shape does not fit in shared memory");
    g.stages = (int)((budget - off) / g.stage_bytes);
    if (g.stages > kMaxTmStages) g.stages = kMaxTmStages;
    size_t smem = (size_t)off + (size_t)g.stages * g.stage_bytes + 1024 + slack;
    if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM (each allocates all of TMEM)
    MC_CHECK(smem <= 226 * 1024, "token_mix: shape does not fit in shared memory");
    MC_CHECK(mode != TM_FWD || p->x != p->y, "token_mix fwd: x and y must not alias");

    CUtensorMap tmU, tmDY, tmX, tmS;
    memset(&tmDY, 0, sizeof(tmDY));
    memset(&tmX, 0, sizeof(tmX));
    memset(&tmS, 0, sizeof(tmS));
    int rc = tm_make_map(&tmU, p->u, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.D, g.P, g.B, g.D, (int64_t)g.P * g.D, 64, g.Ppad,
                         CU_TENSOR_MAP_SWIZZLE_128B, "u");
    if (rc != MC_OK) return rc;
    if (mode != TM_FWD) {
        MC_CHECK(p->dy != nullptr, "token_mix: dy is null");
        rc = tm_make_map(&tmDY, p->dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.D, g.P, g.B, g.D, (int64_t)g.P * g.D, 64, g.Ppad,
                         CU_TENSOR_MAP_SWIZZLE_128B, "dy");
        if (rc != MC_OK) return rc;
    } else {
        MC_CHECK(p->x != nullptr && p->y != nullptr && p->b2 != nullptr, "token_mix fwd: null operand");
        rc = tm_make_map(&tmX, p->x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.D, g.P, g.B, g.D, (int64_t)g.P * g.D, 32, g.Ppad,
                         CU_TENSOR_MAP_SWIZZLE_NONE, "x");
        if (rc != MC_OK) return rc;
    }
    if (mode == TM_DGRAD) MC_CHECK(p->y != nullptr, "token_mix dgrad: null output");
    if (g.spill) {
        MC_CHECK(p->spill_ld >= g.H, "token_mix: spill pitch too small");
        rc = tm_make_map(&tmS, p->spill, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.H, g.D, g.B, p->spill_ld, (int64_t)g.D * p->spill_ld,
                         64, 128, CU_TENSOR_MAP_SWIZZLE_128B, "spill");
        if (rc != MC_OK) return rc;
    }
    switch (mode) {
        case TM_FWD: return tm_launch<TM_FWD>(tmU, tmDY, tmX, tmS, g, grid, smem, stream);
        case TM_DGRAD: return tm_launch<TM_DGRAD>(tmU, tmDY, tmX, tmS, g, grid, smem, stream);
        default: return tm_launch<TM_WGRAD>(tmU, tmDY, tmX, tmS, g, grid, smem, stream);
    }
}

}  // namespace

}  // namespace mc

using namespace mc;

extern "C" int mc_token_mix_supported(int64_t P, int64_t D) {
    return (P >= 1 && P <= 80 && D >= 128 && D % 128 == 0) ? 1 : 0;
}
extern "C" int mc_token_mix_fwd(const mc_token_mix_params* p, void* stream) {
    return tm_run(p, TM_FWD, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int mc_token_mix_dgrad(const mc_token_mix_params* p, void* stream) {
    return tm_run(p, TM_DGRAD, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int mc_token_mix_wgrad(const mc_token_mix_params* p, void* stream) {
    return tm_run(p, TM_WGRAD, reinterpret_cast<cudaStream_t>(stream));
}
