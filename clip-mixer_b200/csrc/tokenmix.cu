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
//   warps 0-7   E2: TMEM -> + residual + bias -> global (forward) / plain store (dgrad)
//   warps 8-23  E1: TMEM -> bias + QuickGELU (or QuickGELU' product) -> bf16 -> smem operand tile (32-column chunks
//               dealt round-robin to four groups of four warps, one warp per TMEM lane quarter)
//   warp 24     TMA producer: activation tiles (bf16 [P x 128] as two 64-wide swizzled groups) + L2 prefetch of the
//               fp32 residual tile
//   warp 25     TMEM allocator, barrier init
//   warp 26     MMA issuer of the "down" GEMMs      warp 27   MMA issuer of the "up" GEMMs   (one elected lane each)
//   WGRAD mode: warps 0-7 idle; the weight-gradient accumulators stay in TMEM across all tiles of the CTA.
// Role order matters: the warp scheduler favours the highest warp id of an SM sub-partition, the MMAs here are tiny
// (32-64 clocks each), so the kernel is paced by how fast ONE thread gets its tcgen05.mma / commit / try_wait
// instructions issued (measured with the -DTM_TRACE timeline: with the MMA issuer as warp 1 below sixteen polling
// epilogue warps it needed ~1000 clocks per four MMAs).  Waiting epilogue warps back off with nanosleep.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"

namespace mc {

namespace {

constexpr int kE1Warps = 16;                                   // four per TMEM lane quarter
constexpr int kE2Warps = 8;
constexpr int kLnRows = 10;                                    // token rows per E2 warp in the fused-LayerNorm prologue (P <= 80)
constexpr int kE1Groups = kE1Warps / 4;
constexpr int kTmThreads = 128 + 32 * (kE1Warps + kE2Warps);   // 768
constexpr int kProdWarp = kE1Warps + kE2Warps, kAllocWarp = kProdWarp + 1, kInitWarp = kAllocWarp, kMmaDnWarp = kProdWarp + 2,
              kMmaWarp = kProdWarp + 3;
constexpr int kMaxAtoms = 5;                                   // hidden width 4P <= 320
constexpr int kMaxTmStages = 2;
constexpr uint32_t kAtomBytes = 128 * 128;                     // [128 rows x 64 bf16] swizzled K-major atom

enum { TM_FWD = 0, TM_DGRAD = 1, TM_WGRAD = 2 };

// Debug timeline (build with -DTM_TRACE, see tools/tokenmix_trace.sh): CTA 0 records (tag, clock) per role.
#ifdef TM_TRACE
__device__ unsigned long long g_tm_trace[28][1024];   // indexed by warp
#define TM_TR(role, tag)                                                                                   \
    do {                                                                                                   \
        if (blockIdx.x == 0 && lane == 0 && tr_ctr < 1024)                                                 \
            g_tm_trace[role][tr_ctr++] = ((unsigned long long)(tag) << 48) | ((unsigned long long)clock64() & 0xffffffffffffull); \
    } while (0)
#else
#define TM_TR(role, tag) do { } while (0)
#endif

struct TmArgs {
    int B, P, D, H;
    int Ppad, Hpad, natoms;
    int tiles_d, num_tiles;
    int SW;                   // hidden columns per segment (multiple of 64)
    int nbuf;                 // working TMEM buffers: 2 = segments ping-pong and the up GEMMs of a tile overlap the E1 pass of
                              // the previous one; 1 = one buffer (backward: Z and dH side by side leave no room for two)
    int zpitch;               // TMEM columns between Z and dH inside a working buffer (>= widest segment, multiple of 32)
    int ybufs;                // output tiles in TMEM (2 forward, 1 dgrad)
    int stages;
    // FWD with LayerNorm in the prologue (ln_sums != nullptr): the E2 warps read the fp32 block input themselves, normalise it
    // with the row statistics the producing GEMM left behind (sum, sum of squares per token row) and write the bf16 operand
    // tile straight into the swizzled smem stage (and to u_out for the backward kernels); no TMA load of u, no LayerNorm kernel
    const float* ln_sums;     // [B*P][2]
    const float* ln_gamma;
    const float* ln_beta;
    __nv_bfloat16* u_out;     // [B, P, D]
    float* ln_mean;           // [B*P] out (LayerNorm backward reads them)
    float* ln_rstd;
    int early;                // FWD: an E1 warp hands the Z buffer back as soon as its last chunk of the segment is in registers (before
                              // the QuickGELU / store work), so the up GEMMs of the segment after next start ~1 k clocks earlier
    int stagger;              // FWD, two Z buffers: E1 groups {0,1} own the segments of buffer 0, groups {2,3} those of buffer 1 (two
                              // chunks per warp and segment), so the two pairs are in different phases instead of all 16 warps
                              // waiting, computing and storing in lock-step
    int aug;                  // 1: lin1's bias rides in the up GEMM (needs two spare k-rows, Ppad - P >= 2): U rows P, P+1 are
                              // constant ones, W1^T rows P, P+1 hold the bf16 hi / lo parts of b1 - the E1 warps add nothing
    uint32_t u_tx_bytes;      // bytes one TMA load of a U group delivers (aug: P rows, else Ppad rows)
    uint32_t grp_bytes;       // Ppad * 128: one [Ppad x 64] swizzled group of an activation / weight tile
    uint32_t a_grp_bytes;     // group pitch of the U tile in smem (>= grp_bytes; WGRAD: 16 KB, rows up to 128)
    uint32_t off_w1t, off_w2t, off_h, off_h2, off_stage, stage_bytes, off_b1, off_b2;
    const __nv_bfloat16* w1;
    int ld1;
    const __nv_bfloat16* w2;
    int ld2;
    const float* b1;
    const float* b2;
    const float* x;
    float* y;
    // WGRAD: the hidden units are dealt to CTAs in slices of slice_w columns
    int nslices, slice_w;
    int slice_cta0[5];        // CTAs [slice_cta0[s], slice_cta0[s+1]) work on hidden slice s (wider slices get more CTAs)
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
__device__ __forceinline__ bool tm_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tm_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tm_tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// tanh.approx.f32 runs at the full MUFU rate (16 lanes/clk/SM, measured with tools/ubench/mufu_rate.cu; the packed
// f16x2 form is split into two MUFU ops and gains nothing).  sigmoid(1.702 z) = 0.5 + 0.5 tanh(0.851 z) costs one
// MUFU op per element where ex2 + rcp costs two.
#ifdef TM_FAKE_TANH   // experiment only (wrong results): how fast is the kernel without the MUFU work?
__device__ __forceinline__ float2 tm_tanh2(float2 a) { return __fmul2_rn(a, make_float2(0.25f, 0.25f)); }
#else
__device__ __forceinline__ float2 tm_tanh2(float2 a) { return make_float2(tanh_approx(a.x), tanh_approx(a.y)); }
#endif
// QuickGELU (model.py:175-177) on packed pairs: x * sigmoid(1.702 x) = hx + hx * tanh(0.851 x)
__device__ __forceinline__ float2 tm_gelu2(float2 x) {
    const float2 t = tm_tanh2(__fmul2_rn(x, make_float2(0.851f, 0.851f)));
    const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    return __ffma2_rn(hx, t, hx);
}
// sigmoid(1.702 z) = 0.5 + 0.5 tanh(0.851 z)
__device__ __forceinline__ float2 tm_sigmoid2(float2 z) {
    const float2 t = tm_tanh2(__fmul2_rn(z, make_float2(0.851f, 0.851f)));
    return __ffma2_rn(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
}
__device__ __forceinline__ float2 tm_gelu_grad_from_s(float2 z, float2 s) {
    const float2 w = __fmul2_rn(z, make_float2(kGeluA, kGeluA));
    const float2 nw = __fmul2_rn(z, make_float2(-kGeluA, -kGeluA));
    const float2 t1 = __ffma2_rn(nw, s, w);      // w (1 - s)
    return __ffma2_rn(t1, s, s);                 // s + w s (1 - s)
}

// Resident weight tiles.  Element (p, jl) of a tile lives at
//   (jl / 64) * grp_bytes + p * 128 + (((jl % 64) / 8) ^ (p % 8)) * 16 + (jl % 8) * 2
// which is exactly what TMA's SWIZZLE_128B produces for a [Ppad x 64] box of a [P x 4P] row-major matrix, so both tiles
// arrive by TMA tensor loads in the prologue (w2t from W2, w1t from the W1^T copy the host refreshes once per step with
// mc_transpose_bf16); rows >= P and columns >= 4P are zero-filled by TMA.  Round 1 gathered W1^T with 2-byte loads and
// swizzled both tiles through registers in every launch: 13 k clocks of a 71 k-clock kernel (-DTM_TRACE timeline).
// MODE: TM_FWD / TM_DGRAD / TM_WGRAD.  TD: compile-time channel count D (0 = run time) - with it every global access of
// the E2 warps is a single LDG / STG with an immediate offset.
template <int MODE, int TD>
__global__ void __launch_bounds__(kTmThreads, 1)
token_mix_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmDY,
                 const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1T,
                 const __grid_constant__ CUtensorMap tmW2, const TmArgs g) {
    extern __shared__ uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t u_full[kMaxTmStages];
    __shared__ __align__(8) uint64_t u_empty[kMaxTmStages];
    __shared__ __align__(8) uint64_t z_full[2];
    __shared__ __align__(8) uint64_t z_empty[2];
    __shared__ __align__(8) uint64_t h_full[kMaxAtoms];
    __shared__ __align__(8) uint64_t h_empty[kMaxAtoms];
    __shared__ __align__(8) uint64_t y_full[2];
    __shared__ __align__(8) uint64_t y_empty[2];
    __shared__ __align__(8) uint64_t w_full;
    __shared__ uint32_t tmem_base_smem;

#ifndef MC_DIVERGENT_WARP_IDX
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform role index (see gemm_tc.cu)
#else
    const int warp = threadIdx.x >> 5;
#endif
    const int lane = threadIdx.x & 31;
    [[maybe_unused]] int tr_ctr = 0;
    if (warp == kInitWarp) TM_TR(kInitWarp, 0);
    const uint32_t base = (smem_u32(dyn_smem) + 1023u) & ~1023u;
    const uint32_t w1t = base + g.off_w1t, w2t = base + g.off_w2t, hbuf = base + g.off_h, h2buf = base + g.off_h2;
    float* b1s = reinterpret_cast<float*>(dyn_smem + (base - smem_u32(dyn_smem)) + g.off_b1);
    float* b2s = reinterpret_cast<float*>(dyn_smem + (base - smem_u32(dyn_smem)) + g.off_b2);
    const int SW = g.SW;      // hidden columns per pipeline segment (multiple of 64)
    // WGRAD: CTAs are dealt round-robin to the hidden slices; every slice walks over all tiles
    int slice = 0;
    if (MODE == TM_WGRAD) {
#pragma unroll
        for (int q = 1; q < 4; ++q)
            if (q < g.nslices && (int)blockIdx.x >= g.slice_cta0[q]) slice = q;
    }
    const int jbase = MODE == TM_WGRAD ? slice * g.slice_w : 0;
    const int work0 = MODE == TM_WGRAD ? (int)blockIdx.x - g.slice_cta0[slice] : (int)blockIdx.x;
    const int work_stride = MODE == TM_WGRAD ? g.slice_cta0[slice + 1] - g.slice_cta0[slice] : (int)gridDim.x;
    // hidden columns handled by this CTA (tile-local: column c <-> hidden unit jbase + c)
    int my_hpad = g.Hpad;
    if (MODE == TM_WGRAD) my_hpad = g.Hpad - jbase < g.slice_w ? g.Hpad - jbase : g.slice_w;
    const int my_atoms = (my_hpad + 63) / 64;
    const int nseg = (my_hpad + SW - 1) / SW;

    // ---- prologue: barriers + TMEM, then the resident weights and the first activation tiles are requested through TMA
    //      (one DRAM / L2 round trip for everything); the MMA issuers wait for the weights on w_full ----
    if (warp == kProdWarp && lane == 0) {
        tma_prefetch_desc(&tmU);
        tma_prefetch_desc(&tmW1T);
        tma_prefetch_desc(&tmW2);
        if (MODE != TM_FWD) tma_prefetch_desc(&tmDY);
        if (MODE == TM_FWD) tma_prefetch_desc(&tmX);
    }
    if (warp == kInitWarp && lane == 0) {
        mbar_init(smem_u32(&w_full), 1);
        const bool fuse_ln_init = MODE == TM_FWD && g.ln_sums != nullptr;
        for (int s = 0; s < kMaxTmStages; ++s) {
            mbar_init(smem_u32(&u_full[s]), fuse_ln_init ? kE2Warps : 1);   // fused LayerNorm: the E2 warps fill the stage
            mbar_init(smem_u32(&u_empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&z_full[s]), 1);
            mbar_init(smem_u32(&z_empty[s]), g.stagger ? kE1Warps / 2 : kE1Warps);
            mbar_init(smem_u32(&y_full[s]), 1);
            mbar_init(smem_u32(&y_empty[s]), kE2Warps);
        }
        for (int sg = 0; sg < kMaxAtoms; ++sg) {
            // h_full[sg] / h_empty[sg] hand the bf16 operand atoms of hidden SEGMENT sg between the E1 warps and the
            // down-GEMM issuer: one arrival per valid 32-column chunk and lane quarter
            int c_lo = sg * (SW / 32), c_hi = c_lo + SW / 32, nvalid = 0;
            for (int c = c_lo; c < c_hi; ++c) nvalid += (c * 32 < my_hpad) ? 1 : 0;
            mbar_init(smem_u32(&h_full[sg]), nvalid > 0 ? 4 * nvalid : 1);
            mbar_init(smem_u32(&h_empty[sg]), 1);
        }
        fence_mbar_init();
    }
    if (warp == kAllocWarp) {
        tmem_alloc(smem_u32(&tmem_base_smem), 512);
        tmem_relinquish();
    }
    __syncthreads();
    MC_PDL_PROLOGUE();      // first global-memory access is below (see common.cuh)
    float b1v[2], b2v = 0.f;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int i = threadIdx.x + it * blockDim.x;
        b1v[it] = (i < g.natoms * 64 && jbase + i < g.H) ? g.b1[jbase + i] : 0.f;
    }
    if (MODE == TM_FWD && (int)threadIdx.x < g.P) b2v = g.b2[threadIdx.x];
    // activation tile loads: also used by the producer loop below
    auto issue_tile = [&](int t, uint32_t st) {
        const int b = t / g.tiles_d, d0 = (t - b * g.tiles_d) * 128;
        const uint32_t bar = smem_u32(&u_full[st]);
        const uint32_t dst = base + g.off_stage + st * g.stage_bytes;
        mbar_arrive_expect_tx(bar, 2u * g.u_tx_bytes + (MODE == TM_FWD ? 0u : 2u * g.grp_bytes));
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
    };
    const bool fuse_ln = MODE == TM_FWD && g.ln_sums != nullptr;
    int npre = 0;      // tiles requested in the prologue (one per smem stage)
    if (warp == kProdWarp) {
        if (lane == 0) {
            // resident weights: my_atoms groups of [Ppad x 64] per matrix, hidden columns jbase + 64 a ..
            const uint32_t wb = smem_u32(&w_full);
            mbar_arrive_expect_tx(wb, 2u * (uint32_t)my_atoms * g.grp_bytes);
            for (int a = 0; a < my_atoms; ++a) {
                tma_load_3d(w1t + (uint32_t)a * g.grp_bytes, &tmW1T, wb, jbase + 64 * a, 0, 0);
                tma_load_3d(w2t + (uint32_t)a * g.grp_bytes, &tmW2, wb, jbase + 64 * a, 0, 0);
            }
        }
        if (!fuse_ln) {
            for (int t = work0; t < g.num_tiles && npre < g.stages; t += work_stride, ++npre)
                if (lane == 0) issue_tile(t, (uint32_t)npre);
        }
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int i = threadIdx.x + it * blockDim.x;
        if (i < g.natoms * 64) b1s[i] = b1v[it];
    }
    if (MODE == TM_FWD && (int)threadIdx.x < g.Ppad) b2s[threadIdx.x] = b2v;
    if (MODE == TM_WGRAD || g.aug || fuse_ln) {
        // Constant rows of every U group, written once (TMA only rewrites the rows of its box: P rows with aug, else Ppad):
        //   aug:   rows P and P+1 are all ones - with W1^T rows P / P+1 = hi / lo halves of b1 the up GEMM adds the bias;
        //   WGRAD: row Ppad is all ones (its accumulator lane collects db1 = sum_d dZ1); every other row >= P is zero.
        const uint32_t stage0 = base + g.off_stage;
        const int r_lo = (g.aug || fuse_ln) ? g.P : g.Ppad, r_hi = MODE == TM_WGRAD ? 128 : g.Ppad;   // fused LN: nobody else writes rows >= P
        const int rows_c = r_hi - r_lo;
        for (int s = 0; s < g.stages; ++s)
            for (int grp = 0; grp < 2; ++grp)
                for (int i = threadIdx.x; i < rows_c * 8; i += blockDim.x) {
                    const int r = r_lo + (i >> 3);
                    const bool one = (g.aug && (r == g.P || r == g.P + 1)) || (MODE == TM_WGRAD && r == g.Ppad);
                    const uint32_t v = one ? 0x3f803f80u : 0u;     // bf16 1.0 pairs
                    tm_sts128(stage0 + s * g.stage_bytes + grp * g.a_grp_bytes + r * 128 + ((i & 7) << 4), v, v, v, v);
                }
    }
    if (g.aug) {
        // bias rows of the resident W1^T tile (after its TMA load has landed; rows >= P arrived as zeros)
        mbar_wait(smem_u32(&w_full), 0u);
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int jl = threadIdx.x + it * blockDim.x;
            if (jl < my_atoms * 64) {
                const __nv_bfloat16 hi = __float2bfloat16_rn(b1v[it]);
                const __nv_bfloat16 lo = __float2bfloat16_rn(b1v[it] - __bfloat162float(hi));
                const uint32_t col = (uint32_t)(jl >> 6) * g.grp_bytes + (uint32_t)(jl & 7) * 2u;
                const uint32_t c8 = (uint32_t)((jl & 63) >> 3);
                const uint32_t p0 = (uint32_t)g.P, p1 = (uint32_t)g.P + 1u;
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(w1t + col + p0 * 128u + ((c8 ^ (p0 & 7u)) << 4)),
                             "h"(*reinterpret_cast<const unsigned short*>(&hi)) : "memory");
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(w1t + col + p1 * 128u + ((c8 ^ (p1 & 7u)) << 4)),
                             "h"(*reinterpret_cast<const unsigned short*>(&lo)) : "memory");
            }
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (warp == kInitWarp) TM_TR(kInitWarp, 1);
    // TMEM column map: two rotating working buffers (FWD: Z; else Z and dH side by side), then the output tile(s)
    const uint32_t zpitch = (uint32_t)g.zpitch;
    const uint32_t wbuf_cols = MODE == TM_FWD ? zpitch : 2u * zpitch;
    const uint32_t nbuf = (uint32_t)g.nbuf, ybufs = (uint32_t)g.ybufs;
    const uint32_t col_y = nbuf * wbuf_cols;                                            // FWD: Y0, Y1; DGRAD: dU
    const uint32_t col_acc2 = nbuf * wbuf_cols, col_acc1 = nbuf * wbuf_cols + g.slice_w;   // WGRAD: dW2 / dW1^T accumulators

    if (warp == kProdWarp) {
        // ===================== TMA producer =====================
        if (lane == 0 && !fuse_ln) {
            uint32_t st = 0, ph = 0;
            int it = 0;
            for (int t = work0; t < g.num_tiles; t += work_stride, ++it) {
                if (it >= npre) {      // the first tiles were requested in the prologue
                    mbar_wait_relaxed(smem_u32(&u_empty[st]), ph ^ 1u, 32);
                    TM_TR(kProdWarp, 1);
                    issue_tile(t, st);
                }
                if (++st == (uint32_t)g.stages) {
                    st = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer, "up" GEMMs: Z^T = U^T W1^T (and dH^T = dY^T W2) =====================
        // Two issuing warps (this one and kMmaDnWarp) because the kernel is paced by the serial instruction stream of
        // the issuer: every tcgen05.mma holds the issue slot for >= 64 clocks whatever its N and every barrier wait costs
        // a few hundred (-DTM_TRACE timeline: one issuer needed ~5700 clocks per tile for 21 MMAs and 9 waits).  A
        // tcgen05.commit only tracks the MMAs of its own thread, which is exactly the split needed here.
        // The whole warp runs the loop (all values warp-uniform -> uniform registers); one elected lane issues.
        {
            const bool leader = tm_elect_one();
            uint32_t st = 0, ph = 0, sc = 0;
            const int ksteps_up = g.Ppad / 16;    // 1 .. 5
            // descriptor templates: the start-address field (bits 0-13, address >> 4) is added per instruction
            const uint64_t dMNg = make_sdesc_sw128(0u, g.grp_bytes, 1024u);       // MN-major, 64-wide groups grp_bytes apart
            const uint64_t dMNa = make_sdesc_sw128(0u, g.a_grp_bytes, 1024u);     // MN-major U tile
            mbar_wait(smem_u32(&w_full), 0u);                                     // resident weight tiles have landed
            for (int t = work0; t < g.num_tiles; t += work_stride) {
                TM_TR(kMmaWarp, 1);
                mbar_wait(smem_u32(&u_full[st]), ph);
                TM_TR(kMmaWarp, 2);
                tc_fence_after();
                const uint32_t u_base = base + g.off_stage + st * g.stage_bytes;
                const uint32_t dy_base = u_base + 2 * g.a_grp_bytes;
                const uint64_t ad_u = dMNa + (u_base >> 4), ad_dy = dMNg + (dy_base >> 4);
                for (int s = 0; s < nseg; ++s) {
                    const uint32_t b = nbuf == 2 ? (sc & 1u) : 0u;
                    const uint32_t zph = nbuf == 2 ? ((sc >> 1) & 1u) : (sc & 1u);
                    mbar_wait(smem_u32(&z_empty[b]), zph ^ 1u);
                    TM_TR(kMmaWarp, 3);
                    tc_fence_after();
                    const int c0 = s * SW;
                    const int w = my_hpad - c0 < SW ? my_hpad - c0 : SW;
                    const uint32_t idesc_up = make_idesc_bf16(128, w, 1, 1);
                    const uint32_t goff = (uint32_t)(c0 >> 6) * g.grp_bytes;
                    const uint64_t bd_1 = dMNg + ((w1t + goff) >> 4), bd_2 = dMNg + ((w2t + goff) >> 4);
                    const uint32_t zcol = tmem_base + b * wbuf_cols;
                    if (leader) {
                        // k-step = 16 token rows of the [Ppad x 64] groups = 2048 B (+128 in the descriptor)
#pragma unroll
                        for (int ks = 0; ks < 5; ++ks) {
                            if (ks < ksteps_up) {
                                umma_ss(zcol, ad_u + 128 * ks, bd_1 + 128 * ks, idesc_up, ks > 0 ? 1u : 0u);
                                if (MODE != TM_FWD) umma_ss(zcol + zpitch, ad_dy + 128 * ks, bd_2 + 128 * ks, idesc_up, ks > 0 ? 1u : 0u);
                            }
                        }
                        if (MODE != TM_WGRAD && s == nseg - 1) umma_commit(smem_u32(&u_empty[st]));
                        umma_commit(smem_u32(&z_full[b]));
                    }
                    ++sc;
                }
                if (++st == (uint32_t)g.stages) {
                    st = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == kMmaDnWarp) {
        // ===================== MMA issuer, "down" GEMMs: consume the bf16 atoms the E1 warps produce =====================
        {
            const bool leader = tm_elect_one();
            uint32_t st = 0, n = 0;
            const uint32_t idesc_dn = make_idesc_bf16(128, g.Ppad, 0, 0);
            const uint32_t wdn = MODE == TM_FWD ? w2t : w1t;
            const uint64_t dK = make_sdesc_sw128(0u, 16u, 1024u);                 // K-major operand
            const uint64_t dMNh = make_sdesc_sw128(0u, kAtomBytes, 1024u);        // MN-major view of the H^T / dZ1^T atoms
            mbar_wait(smem_u32(&w_full), 0u);
            for (int t = work0; t < g.num_tiles; t += work_stride, ++n) {
                if (MODE == TM_WGRAD) {
                    // dW2[p, j] += sum_d dY[p, d] H^T[d, j];  dW1^T[p, j] += sum_d U[p, d] dZ1^T[d, j]   (K = 128 channels);
                    // one MMA covers the whole hidden slice (N = up to 128 columns = 2 atoms)
                    const uint32_t u_base = base + g.off_stage + st * g.stage_bytes;
                    const uint32_t dy_base = u_base + 2 * g.a_grp_bytes;
                    for (int sg = 0; sg < nseg; ++sg) mbar_wait(smem_u32(&h_full[sg]), n & 1u);   // all segments of the slice
                    TM_TR(kMmaDnWarp, 5);
                    tc_fence_after();
                    const uint32_t idesc_w = make_idesc_bf16(128, my_hpad, 0, 1);
                    // A (K-major): rows p, 64-channel atoms = the activation groups; k-step 16 channels = 32 B (+2)
                    // B (MN-major): K = channel rows of the [128 d x 64 j] atoms, 16 rows = 2048 B (+128); atoms kAtomBytes apart
                    const uint64_t b_h = dMNh + (hbuf >> 4), b_dz = dMNh + (h2buf >> 4);
                    const uint32_t acc2 = tmem_base + col_acc2, acc1 = tmem_base + col_acc1;
                    const uint32_t accum0 = n > 0 ? 1u : 0u;
                    if (leader) {
#pragma unroll
                        for (int hlf = 0; hlf < 2; ++hlf) {
                            const uint64_t a_dy = dK + ((dy_base + hlf * g.grp_bytes) >> 4);
                            const uint64_t a_u = dK + ((u_base + hlf * g.a_grp_bytes) >> 4);
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                const int ks = hlf * 4 + k4;
                                umma_ss(acc2, a_dy + 2 * k4, b_h + 128 * ks, idesc_w, ks > 0 ? 1u : accum0);
                                umma_ss(acc1, a_u + 2 * k4, b_dz + 128 * ks, idesc_w, ks > 0 ? 1u : accum0);
                            }
                        }
                        for (int sg = 0; sg < nseg; ++sg) umma_commit(smem_u32(&h_empty[sg]));
                        umma_commit(smem_u32(&u_empty[st]));     // the activation tile is dead once these MMAs have read it
                    }
                    if (++st == (uint32_t)g.stages) st = 0;
                } else {
                    const uint32_t yb = ybufs == 2 ? (n & 1u) : 0u;
                    mbar_wait(smem_u32(&y_empty[yb]), (ybufs == 2 ? ((n >> 1) & 1u) : (n & 1u)) ^ 1u);
                    TM_TR(kMmaDnWarp, 4);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + col_y + yb * g.Ppad;
                    for (int sg = 0; sg < nseg; ++sg) {
                        // every barrier wait and every tcgen05.commit costs this thread about as much as an MMA: the
                        // hand-over is per hidden segment (up to 16 k-steps behind one wait), not per 64-column atom
                        mbar_wait(smem_u32(&h_full[sg]), n & 1u);
                        TM_TR(kMmaDnWarp, 5);
                        tc_fence_after();
                        const int k0 = sg * (SW / 16);
                        int k1 = k0 + SW / 16;
                        if (k1 > my_hpad / 16) k1 = my_hpad / 16;
                        if (leader) {
                            for (int ks = k0; ks < k1; ++ks) {
                                // k-step ks: atom ks / 4 (16 KB apart in the H buffer, grp_bytes apart in the weight tile), 32 B inside
                                const uint64_t ad = dK + ((hbuf + (uint32_t)(ks >> 2) * kAtomBytes + (uint32_t)(ks & 3) * 32u) >> 4);
                                const uint64_t bd = dK + ((wdn + (uint32_t)(ks >> 2) * g.grp_bytes + (uint32_t)(ks & 3) * 32u) >> 4);
                                umma_ss(d_tmem, ad, bd, idesc_dn, ks > 0 ? 1u : 0u);
                            }
                            umma_commit(smem_u32(&h_empty[sg]));
                            if (sg == nseg - 1) umma_commit(smem_u32(&y_full[yb]));
                        }
                    }
                }
            }
            if (MODE == TM_WGRAD && leader) umma_commit(smem_u32(&y_full[0]));   // accumulators final
        }
    } else if (warp >= kE2Warps && warp < kE2Warps + kE1Warps) {
        // ===================== E1: hidden activation TMEM -> bf16 smem operand =====================
        const int e = warp - kE2Warps;
        const int q = e & 3, grp = e >> 2;                   // 32-column chunk c of a tile belongs to group c % kE1Groups
        const int r = q * 32 + lane;                         // tile row = channel d0 + r
        const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        uint32_t sc = 0, n = 0;
        for (int t = work0; t < g.num_tiles; t += work_stride, ++n) {
            for (int s = 0; s < nseg; ++s, ++sc) {
                const uint32_t b = nbuf == 2 ? (sc & 1u) : 0u;
                if (g.stagger && b != (uint32_t)(grp >> 1)) continue;      // the other pair of groups owns this segment
                TM_TR(warp, 1);
                mbar_wait_relaxed(smem_u32(&z_full[b]), nbuf == 2 ? ((sc >> 1) & 1u) : (sc & 1u), 20);
                TM_TR(warp, 2);
                tc_fence_after();
                const uint32_t zcol = t_lane + b * wbuf_cols;
                const int a0 = (s * SW) >> 6;
                int a1 = ((s + 1) * SW) >> 6;
                if (a1 > my_atoms) a1 = my_atoms;
                // the down GEMMs of the previous tile have released this segment's operand atoms
                mbar_wait_relaxed(smem_u32(&h_empty[s]), (n & 1u) ^ 1u, 20);
                TM_TR(warp, 3);
                bool released = false;
                for (int c = 2 * a0; c < 2 * a1; ++c) {
                    if (g.stagger ? (c & 1) != (grp & 1) : c % kE1Groups != grp) continue;
                    const int a = c >> 1, par = c & 1;
                    const int col = c * 32;                  // tile-local hidden column of this warp's chunk
                    const int rel = col - s * SW;
                    if (col < my_hpad) {
                        const uint32_t dst = hbuf + a * kAtomBytes + row_off;
                        if (MODE == TM_FWD) {
                            // two 16-column halves: the TMEM load of the second half is in flight while the first one is
                            // turned into bf16 (the MUFU-bound part), so only one load latency per chunk is exposed
                            uint32_t va[16], vb[16];
                            tmem_ld16(zcol + rel, va);
                            if (g.early) {
                                // both halves requested at once; when they have arrived and this is the warp's last chunk of
                                // the segment, the Z buffer goes back to the up-GEMM issuer before any math is done on it
                                tmem_ld16(zcol + rel + 16, vb);
                                tmem_ld_wait();
                                const int cstep = g.stagger ? 2 : kE1Groups;
                                if (c + cstep >= 2 * a1 || (c + cstep) * 32 >= my_hpad) {
                                    tc_fence_before();
                                    __syncwarp();
                                    if (lane == 0) mbar_arrive(smem_u32(&z_empty[b]));
                                    released = true;
                                }
                            } else {
                                tmem_ld_wait();
                                tmem_ld16(zcol + rel + 16, vb);
                            }
                            auto half = [&](const uint32_t (&v)[16], int hf) {
                                uint32_t o[8];
                                if (g.aug) {        // the accumulator already holds W1 u + b1
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const float2 h = tm_gelu2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])));
                                        o[i] = pack_bf16x2(h.x, h.y);
                                    }
                                } else {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const float2 bb = *reinterpret_cast<const float2*>(b1s + col + hf * 16 + 2 * i);
                                        const float2 z = __fadd2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), bb);
                                        const float2 h = tm_gelu2(z);
                                        o[i] = pack_bf16x2(h.x, h.y);
                                    }
                                }
                                const uint32_t cc = (uint32_t)(par * 4 + hf * 2);
                                tm_sts128(dst + ((cc ^ sw) << 4), o[0], o[1], o[2], o[3]);
                                tm_sts128(dst + (((cc + 1) ^ sw) << 4), o[4], o[5], o[6], o[7]);
                            };
                            half(va, 0);
                            if (!g.early) tmem_ld_wait();
                            half(vb, 1);
                        } else {
                            const uint32_t dst2 = h2buf + a * kAtomBytes + row_off;
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {
                                uint32_t zv[16], dv[16];
                                tmem_ld16(zcol + rel + hf * 16, zv);
                                tmem_ld16(zcol + zpitch + rel + hf * 16, dv);
                                tmem_ld_wait();
                                uint32_t o[8], oh[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    float2 z = make_float2(__uint_as_float(zv[2 * i]), __uint_as_float(zv[2 * i + 1]));
                                    if (!g.aug) z = __fadd2_rn(z, *reinterpret_cast<const float2*>(b1s + col + hf * 16 + 2 * i));
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
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&h_full[s]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                TM_TR(warp, 4);
                if (lane == 0 && !released) mbar_arrive(smem_u32(&z_empty[b]));
            }
        }
        if (MODE == TM_WGRAD) {
            // ---- final: accumulators -> global gradients.  Lane = token p (or the ones-row Ppad -> db1) ----
            if (e == 0 && lane == 0) pdl_launch_dependents();
            mbar_wait_relaxed(smem_u32(&y_full[0]), 0u, 64);
            tc_fence_after();
            const int p = r;
            for (int c = grp * 32; c < my_hpad; c += 32 * kE1Groups) {
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
    } else if (warp < kE2Warps) {
        // ===================== E2: output accumulator -> global =====================
        if (MODE != TM_WGRAD) {
            const int e2 = warp;
            const int q = e2 & 3, hsel = e2 >> 2;
            const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
            const int nch = g.Ppad / 16;
            const int Dd = TD ? TD : g.D;
            uint32_t n = 0;
            // ---- LayerNorm in the prologue (model.py:216 folded into :220-222): this warp group produces the bf16 operand
            //      tile of tile `tp` in stage j % stages from the fp32 block input and the row sums of the producing GEMM ----
            auto produce = [&](int tp, uint32_t j) {
                const uint32_t st_ = j % (uint32_t)g.stages, use = j / (uint32_t)g.stages;
                const int b = tp / g.tiles_d, d0 = (tp - b * g.tiles_d) * 128;
                // warp e2 owns tokens e2, e2 + 8, ...; lane = four consecutive channels: a token row of the slab is one
                // 512-byte line of loads, one 256-byte line of bf16 stores
                const int c0 = 4 * lane;
                TM_TR(warp, 4);
                const long long row0 = (long long)b * g.P + e2;
                const float* xg = g.x + row0 * Dd + d0 + c0;
                float4 xv[kLnRows];
#pragma unroll
                for (int k = 0; k < kLnRows; ++k)
                    if (e2 + 8 * k < g.P) xv[k] = __ldg(reinterpret_cast<const float4*>(xg + (long long)(8 * k) * Dd));
                // lane k holds the statistics of this warp's k-th token
                float mean_l = 0.f, rstd_l = 0.f;
                if (lane < kLnRows && e2 + 8 * lane < g.P) {
                    const float2 sm = __ldg(reinterpret_cast<const float2*>(g.ln_sums) + row0 + 8 * lane);
                    const float invD = 1.0f / (float)Dd;
                    mean_l = sm.x * invD;
                    rstd_l = rsqrtf(fmaxf(fmaf(-mean_l, mean_l, sm.y * invD), 0.f) + kLnEps);
                    if (d0 == 0) {                                             // one slab per sample writes the saved statistics
                        g.ln_mean[row0 + 8 * lane] = mean_l;
                        g.ln_rstd[row0 + 8 * lane] = rstd_l;
                    }
                }
                const float4 gam = __ldg(reinterpret_cast<const float4*>(g.ln_gamma + d0 + c0));
                const float4 bet = __ldg(reinterpret_cast<const float4*>(g.ln_beta + d0 + c0));
                if (use > 0) mbar_wait_relaxed(smem_u32(&u_empty[st_]), (use - 1u) & 1u, 32);   // the up GEMMs have read the old tile
                TM_TR(warp, 5);
                const uint32_t sbase = base + g.off_stage + st_ * g.stage_bytes + (uint32_t)(lane >> 4) * g.a_grp_bytes +
                                       (uint32_t)(lane & 1) * 8u;
                const uint32_t c8 = (uint32_t)(lane & 15) >> 1;               // 16-byte chunk of the 64-channel group
                __nv_bfloat16* ug = g.u_out + row0 * Dd + d0 + c0;
#pragma unroll
                for (int k = 0; k < kLnRows; ++k) {
                    const int p = e2 + 8 * k;
                    const float mean = __shfl_sync(0xffffffffu, mean_l, k), rstd = __shfl_sync(0xffffffffu, rstd_l, k);
                    if (p < g.P) {
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(fmaf((xv[k].x - mean) * rstd, gam.x, bet.x),
                                                                        fmaf((xv[k].y - mean) * rstd, gam.y, bet.y));
                        const __nv_bfloat162 hi = __floats2bfloat162_rn(fmaf((xv[k].z - mean) * rstd, gam.z, bet.z),
                                                                        fmaf((xv[k].w - mean) * rstd, gam.w, bet.w));
                        uint2 pk;
                        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sbase + (uint32_t)p * 128u + ((c8 ^ ((uint32_t)p & 7u)) << 4)),
                                     "r"(pk.x), "r"(pk.y) : "memory");
                        *reinterpret_cast<uint2*>(ug + (long long)(8 * k) * Dd) = pk;
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&u_full[st_]));
                TM_TR(warp, 6);
            };
            // production runs TWO tiles ahead of the store pass (both stages): the up GEMMs never wait for this warp group
            if (fuse_ln && work0 < g.num_tiles) produce(work0, 0u);
            if (fuse_ln && work0 + work_stride < g.num_tiles) produce(work0 + work_stride, 1u);
            // store pass of one tile.  KCH = 16-token chunks per warp (2 up to 64 tokens, else 3), a compile-time constant,
            // and never more than 32 residual values in flight: with 48 the compiler spilled some right behind their loads,
            // which put a full global-load latency on this warp's critical path in every tile (-DTM_TRACE timeline,
            // profiles/r2h_tokenmix_fwd_trace.txt: tag 6 -> 1 in the fused variant, 3 -> 1 in the other)
            auto store_tile = [&](auto kch_tag, int t) {
                constexpr int KCH = decltype(kch_tag)::value;
                const int b = t / g.tiles_d, d0 = (t - b * g.tiles_d) * 128;
                // this warp owns the 16-token chunks hsel, hsel + 2, hsel + 4: token p = hsel*16 + 32*k + i
                const long long gbase = ((long long)b * g.P + hsel * 16) * Dd + d0 + q * 32 + lane;
                const float* xt = g.x + gbase;
                float* yt = g.y + gbase;
                // residual values: two chunks in flight before the accumulator is waited for; a third chunk (65..80 tokens,
                // warps 0-3 only) reuses the first slot as soon as chunk 0 is stored
                float xr[2][16];
                auto load_chunk = [&](int k, float (&dst)[16]) {
                    const int p0 = hsel * 16 + 32 * k;
                    if (p0 + 16 <= g.P) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) dst[i] = __ldg(xt + (32 * k + i) * Dd);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) dst[i] = p0 + i < g.P ? __ldg(xt + (32 * k + i) * Dd) : 0.f;
                    }
                };
                if (MODE == TM_FWD) {
                    load_chunk(0, xr[0]);
                    load_chunk(1, xr[1]);
                }
                const uint32_t yb = ybufs == 2 ? (n & 1u) : 0u;
                TM_TR(warp, 1);
                mbar_wait_relaxed(smem_u32(&y_full[yb]), ybufs == 2 ? ((n >> 1) & 1u) : (n & 1u), 64);
                TM_TR(warp, 2);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < KCH; ++k) {
                    const int ch = hsel + 2 * k, p0 = hsel * 16 + 32 * k;
                    if (ch < nch) {
                        uint32_t v[16];
                        tmem_ld16(t_lane + col_y + yb * g.Ppad + ch * 16, v);
                        float o[16];
                        if (MODE == TM_FWD) {
#pragma unroll
                            for (int i4 = 0; i4 < 4; ++i4) {
                                const float4 bv = *reinterpret_cast<const float4*>(b2s + p0 + 4 * i4);
                                o[4 * i4] = xr[k & 1][4 * i4] + bv.x;
                                o[4 * i4 + 1] = xr[k & 1][4 * i4 + 1] + bv.y;
                                o[4 * i4 + 2] = xr[k & 1][4 * i4 + 2] + bv.z;
                                o[4 * i4 + 3] = xr[k & 1][4 * i4 + 3] + bv.w;
                            }
                            if (k + 2 < KCH && hsel + 2 * (k + 2) < nch) load_chunk(k + 2, xr[k & 1]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = 0.f;
                        }
                        tmem_ld_wait();
                        if (p0 + 16 <= g.P) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) yt[(32 * k + i) * Dd] = o[i] + __uint_as_float(v[i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (p0 + i < g.P) yt[(32 * k + i) * Dd] = o[i] + __uint_as_float(v[i]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                TM_TR(warp, 3);
                if (lane == 0) mbar_arrive(smem_u32(&y_empty[yb]));
            };
            for (int t = work0; t < g.num_tiles; t += work_stride, ++n) {
                if (fuse_ln) {
                    if (t + 2 * work_stride < g.num_tiles) produce(t + 2 * work_stride, n + 2u);   // its stage: tile t's, read by now
                    if (e2 == 0 && lane == 0 && t + 3 * work_stride < g.num_tiles) {        // pull the tile after that into L2
                        const int t2 = t + 3 * work_stride, b2_ = t2 / g.tiles_d, d2 = (t2 - b2_ * g.tiles_d) * 128;
#pragma unroll
                        for (int i = 0; i < 4; ++i) tm_tma_prefetch_3d(&tmX, d2 + 32 * i, 0, b2_);
                    }
                }
                if (nch <= 4) store_tile(std::integral_constant<int, 2>{}, t);
                else store_tile(std::integral_constant<int, 3>{}, t);
            }
            if (e2 == 0 && lane == 0) pdl_launch_dependents();   // last tile of this CTA stored
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kInitWarp) TM_TR(kInitWarp, 9);
    if (warp == kAllocWarp) tmem_dealloc(tmem_base, 512);
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

template <int MODE, int TD>
int tm_launch_t(const CUtensorMap& tmU, const CUtensorMap& tmDY, const CUtensorMap& tmX, const CUtensorMap& tmW1T,
                const CUtensorMap& tmW2, const TmArgs& g, int grid, size_t smem, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        MC_CUDA(cudaFuncSetAttribute(token_mix_kernel<MODE, TD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        attr_set = true;
    }
    MC_LAUNCH((token_mix_kernel<MODE, TD>), grid, kTmThreads, smem, stream, tmU, tmDY, tmX, tmW1T, tmW2, g);
    MC_CUDA(cudaGetLastError());
#ifdef TM_TRACE
    {
        static unsigned long long host[28][1024];
        cudaStreamSynchronize(stream);
        cudaMemcpyFromSymbol(host, g_tm_trace, sizeof(host));
        fprintf(stderr, "[tm_trace] mode %d P %d D %d\n", MODE, g.P, g.D);
        for (int r = 0; r < 28; ++r)
            for (int i = 0; i < 1024 && host[r][i] != 0; ++i)
                fprintf(stderr, "[tm_trace] role %d ev %d tag %llu clk %llu\n", r, i, host[r][i] >> 48, host[r][i] & 0xffffffffffffull);
        static unsigned long long zero[28][1024];
        cudaMemcpyToSymbol(g_tm_trace, zero, sizeof(zero));
    }
#endif
    return MC_OK;
}

template <int MODE>
int tm_launch(const CUtensorMap& tmU, const CUtensorMap& tmDY, const CUtensorMap& tmX, const CUtensorMap& tmW1T,
              const CUtensorMap& tmW2, const TmArgs& g, int grid, size_t smem, cudaStream_t stream) {
    // the two production widths get their own instantiation (immediate-offset global accesses in the E2 warps)
    if (MODE != TM_WGRAD && g.D == 768) return tm_launch_t<MODE, 768>(tmU, tmDY, tmX, tmW1T, tmW2, g, grid, smem, stream);
    if (MODE != TM_WGRAD && g.D == 512) return tm_launch_t<MODE, 512>(tmU, tmDY, tmX, tmW1T, tmW2, g, grid, smem, stream);
    return tm_launch_t<MODE, 0>(tmU, tmDY, tmX, tmW1T, tmW2, g, grid, smem, stream);
}

int round_up_i(int x, int m) { return (x + m - 1) / m * m; }

int tm_run(const mc_token_mix_params* p, int mode, cudaStream_t stream) {
    MC_CHECK(p != nullptr, "token_mix: null params");
    MC_CHECK(p->B > 0 && p->P > 0 && p->D > 0, "token_mix: empty problem");
    MC_CHECK(mc_token_mix_supported(p->P, p->D), "token_mix: unsupported shape P=%lld D=%lld (need P <= 80, D %% 128 == 0)",
             (long long)p->P, (long long)p->D);
    const bool fuse_ln_h = mode == TM_FWD && p->ln_sums != nullptr;
    MC_CHECK((p->u || fuse_ln_h) && p->w1 && p->w2 && p->b1, "token_mix: null operand");
    MC_CHECK(!fuse_ln_h || (p->ln_gamma && p->ln_beta && p->u_out && p->ln_mean && p->ln_rstd),
             "token_mix fwd with fused LayerNorm: ln_gamma, ln_beta, u_out, ln_mean, ln_rstd are required");
    MC_CHECK(!fuse_ln_h || (((reinterpret_cast<uintptr_t>(p->x) | reinterpret_cast<uintptr_t>(p->ln_gamma) |
                              reinterpret_cast<uintptr_t>(p->ln_beta) | reinterpret_cast<uintptr_t>(p->u_out)) & 15) == 0 &&
                            (reinterpret_cast<uintptr_t>(p->ln_sums) & 7) == 0),
             "token_mix fwd with fused LayerNorm: x, ln_gamma, ln_beta, u_out must be 16-byte aligned, ln_sums 8-byte aligned");
    MC_CHECK(p->w1t != nullptr && p->ld1t >= 4 * p->P && p->ld1t % 8 == 0 && (reinterpret_cast<uintptr_t>(p->w1t) & 15) == 0,
             "token_mix: w1t (W1^T bf16 [P x ld1t], mc_transpose_bf16 of w1; ld1t >= 4P, multiple of 8, 16-byte aligned) is required");
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
    {
        const char* na = getenv("MC_TM_NO_AUG");       // A/B knob: bias added by the E1 warps as in round 1
        g.aug = (g.Ppad - g.P >= 2 && !(na != nullptr && atoi(na) != 0)) ? 1 : 0;
    }
    {
        // both on by default since round 2 (profiles/r2g_tokenmix_stagger_early.txt: forward 37.0 / 36.9 -> 35.6 / 33.1 us for the
        // image / text tower; early release alone is a loss, it needs the staggered groups); MC_TM_STAGGER=0 / MC_TM_EARLY=0 for A/B
        const char* sg = getenv("MC_TM_STAGGER");
        g.stagger = (mode == TM_FWD && !(sg != nullptr && atoi(sg) == 0)) ? 1 : 0;      // needs nbuf == 2 (checked below)
        const char* er = getenv("MC_TM_EARLY");
        g.early = (mode == TM_FWD && !(er != nullptr && atoi(er) == 0)) ? 1 : 0;
    }
    const int u_rows = g.aug ? g.P : g.Ppad;
    g.u_tx_bytes = (uint32_t)u_rows * 128u;
    g.w1 = reinterpret_cast<const __nv_bfloat16*>(p->w1); g.ld1 = (int)p->ld1;
    g.w2 = reinterpret_cast<const __nv_bfloat16*>(p->w2); g.ld2 = (int)p->ld2;
    g.b1 = p->b1; g.b2 = p->b2; g.x = p->x; g.y = p->y;
    g.ln_sums = fuse_ln_h ? p->ln_sums : nullptr;
    g.ln_gamma = p->ln_gamma; g.ln_beta = p->ln_beta; g.u_out = reinterpret_cast<__nv_bfloat16*>(p->u_out);
    g.ln_mean = p->ln_mean; g.ln_rstd = p->ln_rstd;
    MC_CHECK(g.ld1 >= g.P && g.ld2 >= g.H && g.ld1 % 8 == 0 && g.ld2 % 8 == 0, "token_mix: weight pitches must be >= the row length and multiples of 8");
    MC_CHECK((reinterpret_cast<uintptr_t>(p->w1) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w2) & 15) == 0,
             "token_mix: weights must be 16-byte aligned");

    const int sms = sm_count();
    int grid = g.num_tiles < sms ? g.num_tiles : sms;
    {   // balanced persistent grid (same rule as csrc/gemm_tc.cu): only as many CTAs as the round count needs
        const char* bal = getenv("MC_GEMM_BALANCED");
        if (bal != nullptr && atoi(bal) != 0 && g.num_tiles > sms) {
            const int rounds = (g.num_tiles + sms - 1) / sms;
            grid = (g.num_tiles + rounds - 1) / rounds;
        }
    }
    int natoms_smem = g.natoms;
    g.slice_w = 128;
    g.nslices = 1;
    // TMEM plan (512 columns).  A tcgen05.mma costs max(64, N / 2) clocks whatever it computes, so segments are as wide as
    // the columns allow:  forward  two ping-pong Z buffers of 128 + two output tiles;
    //                     dgrad    ONE buffer holding Z and dH of a whole segment (all of 4P when it fits) + one output tile;
    //                     wgrad    Z and dH of the 128-column slice + the two 128-column accumulators.
    if (mode == TM_FWD) {
        g.SW = 128; g.nbuf = 2; g.zpitch = 128; g.ybufs = 2;
    } else if (mode == TM_DGRAD) {
        g.nbuf = 1; g.ybufs = 1;
        const int cap = (512 - g.Ppad) / 2;                         // columns available to each of Z, dH
        g.SW = round_up_i(g.Hpad, 32) <= cap ? round_up_i(g.Hpad, 64) : cap / 64 * 64;
        const int widest = g.Hpad < g.SW ? g.Hpad : g.SW;
        g.zpitch = round_up_i(widest, 32);
    } else {
        g.SW = 128; g.nbuf = 1; g.zpitch = 128; g.ybufs = 1;
    }
    {
        // Experiment knob MC_TM_SW_{FWD,DGRAD,WGRAD}=64|128: narrower segments with TWO working buffers when the columns
        // allow it - the up GEMMs of segment s+1 then overlap the E1 pass of segment s, and with 64-column segments the
        // four E1 groups split into two pairs that work on different segments at the same time instead of in lock-step.
        static const char* names[3] = {"MC_TM_SW_FWD", "MC_TM_SW_DGRAD", "MC_TM_SW_WGRAD"};
        const char* v = getenv(names[mode]);
        const int sw = v ? atoi(v) : 0;
        if (sw == 64 || sw == 128) {
            const int per_buf = (mode == TM_FWD ? 1 : 2) * sw;
            const int fixed = mode == TM_FWD ? 2 * g.Ppad : mode == TM_DGRAD ? g.Ppad : 2 * g.slice_w;
            if ((g.Hpad + sw - 1) / sw <= kMaxAtoms) {
                g.SW = sw;
                g.zpitch = sw;
                g.nbuf = (2 * per_buf + fixed <= 512) ? 2 : 1;
                MC_CHECK(g.nbuf * per_buf + fixed <= 512, "token_mix: MC_TM_SW does not fit in tensor memory");
            }
        }
    }
    MC_CHECK(g.SW >= 64 && g.SW <= 256, "token_mix: bad segment width");
    if (g.nbuf != 2) g.stagger = 0;
    if (!g.stagger) g.early = 0;
    if (mode == TM_WGRAD) {
        g.nslices = (g.Hpad + g.slice_w - 1) / g.slice_w;
        natoms_smem = g.slice_w / 64;
        MC_CHECK(p->gw1 && p->gw2 && p->gb1 && p->dy, "token_mix wgrad: null operand");
        g.gw1 = p->gw1; g.ldg1 = (int)p->ldg1; g.gw2 = p->gw2; g.ldg2 = (int)p->ldg2; g.gb1 = p->gb1;
        MC_CHECK(g.nslices <= 4, "token_mix wgrad: hidden width too large");
        // CTAs per slice in proportion to the slice's cost per tile (the last slice is narrower: 208 = 128 + 80 columns,
        // 320 = 128 + 128 + 64): every slice walks over ALL tiles, so with equal shares the wide slices set the time.
        // Cost per tile ~ fixed hand-over chain + E1 / MMA work proportional to the width (trace: 0.3 + 0.7 w / 128).
        grid = sms;
        if (grid > g.num_tiles * g.nslices) grid = g.num_tiles * g.nslices;
        {
            double cost[4], tot = 0.0;
            for (int q = 0; q < g.nslices; ++q) {
                const int w = g.Hpad - q * g.slice_w < g.slice_w ? g.Hpad - q * g.slice_w : g.slice_w;
                cost[q] = 0.3 + 0.7 * w / 128.0;
                tot += cost[q];
            }
            const char* eq = getenv("MC_TM_WGRAD_EQUAL");
            int used = 0;
            g.slice_cta0[0] = 0;
            for (int q = 0; q < g.nslices; ++q) {
                int n = (eq != nullptr && atoi(eq) != 0) ? grid / g.nslices : (int)(grid * cost[q] / tot + 0.5);
                const int reserve = g.nslices - 1 - q;                    // every later slice needs at least one CTA
                if (q == g.nslices - 1 && !(eq != nullptr && atoi(eq) != 0)) n = grid - used;
                if (n > grid - used - reserve) n = grid - used - reserve;
                if (n > g.num_tiles) n = g.num_tiles;                     // a CTA without a tile would flush an unwritten accumulator
                if (n < 1) n = 1;
                used += n;
                g.slice_cta0[q + 1] = used;
            }
            MC_CHECK(g.slice_cta0[g.nslices] >= g.nslices && g.slice_cta0[g.nslices] <= sms, "token_mix wgrad: bad CTA split");
            grid = g.slice_cta0[g.nslices];
        }
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
    // wgrad keeps the activation tile of the previous work item alive while the next one is loaded
    MC_CHECK(mode != TM_WGRAD || g.stages >= 2, "token_mix wgrad: shape does not fit in shared memory");
    // the fused LayerNorm prologue builds the operand tile two tiles ahead of the store pass
    MC_CHECK(!fuse_ln_h || g.stages >= 2, "token_mix fwd with fused LayerNorm: shape does not fit in shared memory");
    size_t smem = (size_t)off + (size_t)g.stages * g.stage_bytes + 1024 + slack;
    if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM (each allocates all of TMEM)
    MC_CHECK(smem <= 226 * 1024, "token_mix: shape does not fit in shared memory");
    MC_CHECK(mode != TM_FWD || p->x != p->y, "token_mix fwd: x and y must not alias");

    CUtensorMap tmU, tmDY, tmX, tmW1T, tmW2;
    memset(&tmDY, 0, sizeof(tmDY));
    memset(&tmX, 0, sizeof(tmX));
    // resident weight tiles: [Ppad x 64] swizzled groups of W1^T [P x 4P] and W2 [P x 4P]; OOB rows / columns read as zero
    int rcw = tm_make_map(&tmW1T, p->w1t, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.H, g.P, 1, p->ld1t, (int64_t)g.P * p->ld1t, 64,
                          g.Ppad, CU_TENSOR_MAP_SWIZZLE_128B, "w1t");
    if (rcw != MC_OK) return rcw;
    rcw = tm_make_map(&tmW2, p->w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.H, g.P, 1, g.ld2, (int64_t)g.P * g.ld2, 64, g.Ppad,
                      CU_TENSOR_MAP_SWIZZLE_128B, "w2");
    if (rcw != MC_OK) return rcw;
    int rc = tm_make_map(&tmU, fuse_ln_h ? p->u_out : p->u, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.D, g.P, g.B, g.D, (int64_t)g.P * g.D, 64, u_rows,
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
    switch (mode) {
        case TM_FWD: return tm_launch<TM_FWD>(tmU, tmDY, tmX, tmW1T, tmW2, g, grid, smem, stream);
        case TM_DGRAD: return tm_launch<TM_DGRAD>(tmU, tmDY, tmX, tmW1T, tmW2, g, grid, smem, stream);
        default: return tm_launch<TM_WGRAD>(tmU, tmDY, tmX, tmW1T, tmW2, g, grid, smem, stream);
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
