// Micro-benchmark (B200): which part of the GEMM engine's stage hand-over costs the operand delivery rate?
// tools/ubench/l2_ingest shows plain TMA loads landing 114-125 B/clk/SM from L2 through a ring of 32 KB stages released by
// a software mbarrier arrive, while csrc/gemm_tc.cu gets ~42 B/clk/SM through its ring (cta_group::2 loads completing on
// the leader's barrier, stages released by tcgen05.commit, one MMA warp doing wait + fence + 4 MMAs + commit per
// k-block).  This probe adds those pieces one at a time to the same ring (no epilogue; the accumulator is overwritten):
//   single CTA   release = software arrive | tcgen05.commit without MMAs | 4 x tcgen05.mma (128 x BN x 16) + commit
//   CTA pair     2SM TMA loads on the leader's barrier, release = tcgen05.commit multicast without MMAs | with 4 x
//                tcgen05.mma.cta_group::2 (256 x 256 x 16) - the engine's configuration
// Output: clocks per stage and bytes landed per clock and SM; the MMA rows also give the tensor-pipe share (512 or
// 4 * max(64, BN / 2) clocks of MMA work per stage).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ring_handover tools/ubench/ring_handover.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        if (++spins > (1u << 26)) __trap();      // a pipeline bug becomes a CUDA error, not a hung GPU
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t sdesc(uint32_t addr) {   // K-major, 128B swizzle: SBO = 1024 B between 8-row groups
    return uint64_t((addr & 0x3ffffu) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    return el;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int kKB = 16;      // k-blocks per tile (K = 1024)
enum { REL_SOFT = 0, REL_COMMIT = 1, REL_MMA = 2 };

// PAIR: CTA pair (cluster of 2, tcgen05 cta_group::2).  BN: columns of the (pair's) tile.  Stage = A 16 KB + B (BN or
// BN / 2 rows) x 128 B per CTA.
template <bool PAIR, int MODE, int BN, int STAGES>
__global__ void __launch_bounds__(96, 1)
k(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB, int a_blocks, int tiles, long long* clk) {
    extern __shared__ uint8_t dsm[];
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
    __shared__ uint32_t tmem_base_smem;
    constexpr uint32_t kBRows = PAIR ? BN / 2 : BN;
    constexpr uint32_t kStageBytes = 16384 + kBRows * 128;
    const uint32_t base = (smem_u32(dsm) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    uint32_t rank = 0;
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // TMEM: 512 columns (the MMA variants write one accumulator over and over)
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) cluster_sync(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    const long long t0 = clock64();
    uint32_t stage = 0, phase = 0;
    const int work = PAIR ? blockIdx.x / 2 : blockIdx.x, nwork = PAIR ? gridDim.x / 2 : gridDim.x;
    if (warp == 0) {
        // ---- producer: warp-uniform loop, one elected lane issues ----
        const uint32_t el = elect_one();
        for (int t = 0; t < tiles; ++t) {
            const int blk = (int)((work + (long long)t * nwork) % a_blocks) * (PAIR ? 2 : 1) + (int)rank;   // 128-row A block
            for (int kb = 0; kb < kKB; ++kb) {
                mbar_wait(smem_u32(&empty[stage]), phase ^ 1u);
                if (el) {
                    const uint32_t bar = smem_u32(&full[stage]), dst = base + stage * kStageBytes;
                    if (PAIR) {
                        const uint32_t lbar = leader ? bar : mapa(bar, 0);
                        if (leader) mbar_expect(bar, 2u * kStageBytes);
                        tma_load_2d_2sm(dst, &mA, lbar, kb * 64, blk * 128);
                        tma_load_2d_2sm(dst + 16384, &mB, lbar, kb * 64, (int)(rank * kBRows));
                    } else {
                        mbar_expect(bar, kStageBytes);
                        tma_load_2d(dst, &mA, bar, kb * 64, blk * 128);
                        tma_load_2d(dst + 16384, &mB, bar, kb * 64, 0);
                    }
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1 && (!PAIR || leader)) {
        // ---- consumer: waits for the stage, (issues the MMAs,) releases it ----
        const uint32_t el = elect_one();
        constexpr uint32_t id = idesc(PAIR ? 256 : 128, BN);
        for (int t = 0; t < tiles; ++t)
            for (int kb = 0; kb < kKB; ++kb) {
                mbar_wait(smem_u32(&full[stage]), phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_base = base + stage * kStageBytes;
                const uint64_t ad = sdesc(a_base), bd = sdesc(a_base + 16384);
                const uint32_t ebar = smem_u32(&empty[stage]);
                if (el) {
                    if (MODE == REL_SOFT) {
                        mbar_arrive(ebar);
                    } else {
                        if (MODE == REL_MMA) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
                                if (PAIR)
                                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                                 ::"r"(tmem), "l"(ad + 2 * ks), "l"(bd + 2 * ks), "r"(id), "r"(acc) : "memory");
                                else
                                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                                 ::"r"(tmem), "l"(ad + 2 * ks), "l"(bd + 2 * ks), "r"(id), "r"(acc) : "memory");
                            }
                        }
                        if (PAIR)
                            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                         ::"r"(ebar), "h"((uint16_t)3) : "memory");
                        else
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ebar) : "memory");
                    }
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        // drain: the last commits must have landed before TMEM / smem go away
        if (MODE != REL_SOFT) {
            if (el) {
                if (PAIR)
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(smem_u32(&full[0])), "h"((uint16_t)1) : "memory");
                else
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&full[0])) : "memory");
            }
            // full[0] is idle now (every load has been consumed): its next phase completes with this commit
            mbar_wait(smem_u32(&full[0]), ((uint32_t)(tiles * kKB + STAGES - 1) / STAGES) & 1u);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) cluster_sync(); else __syncthreads();
    if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
    if (warp == 2) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Second kernel: the pair pipeline above, dressed step by step like gemm_tc_kernel, to find what takes it from 515 to
// ~690 clocks per k-block (NOT yet run on hardware when committed - round 2 starts here):
//   EPIW    12 extra warps in the engine's layout (epilogue warps 0-11, TMEM allocator 12, B producer 13, A producer 14,
//           issuer 15) that poll an accumulator-full barrier with nanosleep and hand the accumulator back per tile
//   EPIWORK those warps also drain the accumulator (tcgen05.ld, tanh GELU, bf16 pack, swizzled st.shared + proxy fence)
//   WDOG    barrier waits with the engine's globaltimer watchdog slow path
//   RT      stage count / stage size / k-blocks per tile are kernel arguments (no unrolling, runtime multiplies)
//   PROD2   the B operand is issued by a second producer warp
struct RtArgs { int stages, stage_bytes, kb_per_tile; };

__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
template <bool WDOG>
__device__ __forceinline__ void wait2(uint32_t bar, uint32_t parity) {
    if (!WDOG) { mbar_wait(bar, parity); return; }
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    const uint64_t t0 = globaltimer_ns();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (globaltimer_ns() - t0 > 2000000000ull) __trap();
    }
}
__device__ __forceinline__ void wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        __nanosleep(64);
        if (++spins > (1u << 24)) __trap();
    }
}

constexpr int kMaxStages2 = 8;

template <bool EPIW, bool EPIWORK, bool WDOG, bool RT, bool PROD2>
__global__ void __launch_bounds__(512, 1)
k2(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB, int a_blocks, int tiles, long long* clk, RtArgs rt) {
    extern __shared__ uint8_t dsm[];
    __shared__ __align__(8) uint64_t full[kMaxStages2], empty[kMaxStages2], tfull[2], tempty[2];
    __shared__ uint32_t tmem_base_smem;
    const int STAGES = RT ? rt.stages : 5;
    const uint32_t kStageBytes = RT ? (uint32_t)rt.stage_bytes : 32768u;
    const int KB = RT ? rt.kb_per_tile : kKB;
    const uint32_t base = (smem_u32(dsm) + 1023u) & ~1023u;
    const uint32_t epi_base = base + (uint32_t)STAGES * kStageBytes;      // 12 x 4 KB of staging behind the ring
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages2; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tfull[s]), 1);
            mbar_init(smem_u32(&tempty[s]), 24);      // the epilogue warps of BOTH CTAs release the leader's accumulator
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    const long long t0 = clock64();
    const int work = blockIdx.x / 2, nwork = gridDim.x / 2;
    if (warp == 14 || (PROD2 && warp == 13)) {
        const bool do_a = warp == 14, do_b = PROD2 ? warp == 13 : true;
        const uint32_t el = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < tiles; ++t) {
            const int blk = (int)((work + (long long)t * nwork) % a_blocks) * 2 + (int)rank;
            for (int kb = 0; kb < KB; ++kb) {
                wait2<WDOG>(smem_u32(&empty[stage]), phase ^ 1u);
                if (el) {
                    const uint32_t bar = smem_u32(&full[stage]), dst = base + stage * kStageBytes;
                    const uint32_t lbar = leader ? bar : mapa(bar, 0);
                    if (do_a) {
                        if (leader) mbar_expect(bar, 2u * 32768u);
                        tma_load_2d_2sm(dst, &mA, lbar, kb * 64, blk * 128);
                    }
                    if (do_b) tma_load_2d_2sm(dst + 16384, &mB, lbar, kb * 64, (int)(rank * 128));
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 15 && leader) {
        const uint32_t el = elect_one();
        constexpr uint32_t id = idesc(256, 256);
        int stage = 0;
        uint32_t phase = 0, as = 0, aphase = 0;
        for (int t = 0; t < tiles; ++t) {
            if (EPIW) wait2<WDOG>(smem_u32(&tempty[as]), aphase ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem + as * 256;
            for (int kb = 0; kb < KB; ++kb) {
                wait2<WDOG>(smem_u32(&full[stage]), phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_base = base + stage * kStageBytes;
                const uint64_t ad = sdesc(a_base), bd = sdesc(a_base + 16384);
                if (el) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(d_tmem), "l"(ad + 2 * ks), "l"(bd + 2 * ks), "r"(id), "r"(acc) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(smem_u32(&empty[stage])), "h"((uint16_t)3) : "memory");
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            if (el)
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                             ::"r"(smem_u32(&tfull[as])), "h"((uint16_t)3) : "memory");
            as ^= 1u;
            if (as == 0) aphase ^= 1u;
        }
        if (!EPIW) {   // nobody consumes tfull: wait for the last accumulator before tearing down
            const uint32_t last = (uint32_t)(tiles - 1) & 1u, n_done = (uint32_t)(tiles + 1 - (int)last) / 2;   // completions of tfull[last]
            mbar_wait(smem_u32(&tfull[last]), (n_done - 1) & 1u);
        }
    } else if (EPIW && warp < 12) {
        const int q = warp & 3, part = warp >> 2;
        uint32_t as = 0, aphase = 0;
        const uint32_t stage_buf = epi_base + warp * 4096;
        for (int t = 0; t < tiles; ++t) {
            wait_relaxed(smem_u32(&tfull[as]), aphase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (EPIWORK) {
                const uint32_t t_row = tmem + (uint32_t(q * 32) << 16) + as * 256;
                for (int c = part * 32; c < 256; c += 96) {
                    uint32_t v[32];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(t_row + c) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    uint32_t o[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float x0 = __uint_as_float(v[2 * i]), x1 = __uint_as_float(v[2 * i + 1]), t0f, t1f;
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t0f) : "f"(0.851f * x0));
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t1f) : "f"(0.851f * x1));
                        x0 = fmaf(0.5f * x0, t0f, 0.5f * x0);
                        x1 = fmaf(0.5f * x1, t1f, 0.5f * x1);
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[i]) : "f"(x1), "f"(x0));
                    }
                    const uint32_t rowp = stage_buf + lane * 64, sw = (lane >> 1) & 3;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + ((j ^ sw) << 4)), "r"(o[4 * j]), "r"(o[4 * j + 1]),
                                     "r"(o[4 * j + 2]), "r"(o[4 * j + 3]) : "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const uint32_t bar = smem_u32(&tempty[as]);
                if (leader) mbar_arrive(bar);
                else asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(bar, 0)) : "memory");
            }
            as ^= 1u;
            if (as == 0) aphase ^= 1u;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync();
    if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
    if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeTiledFn enc, CUtensorMap* m, void* p, long long rows, int box_rows) {
    cuuint64_t gdim[2] = {1024, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {2048};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    return (int)enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

template <bool PAIR, int MODE, int BN, int STAGES>
void run(const char* name, EncodeTiledFn enc, void* pA, void* pB, long long* clk, int sms, int a_blocks) {
    constexpr int kBRows = PAIR ? BN / 2 : BN;
    constexpr int kStageBytes = 16384 + kBRows * 128;
    CUtensorMap mA, mB;
    if (make_map(enc, &mA, pA, (long long)a_blocks * 128, 128) || make_map(enc, &mB, pB, 256, kBRows)) { printf("encode failed\n"); return; }
    auto fn = k<PAIR, MODE, BN, STAGES>;
    const size_t smem = (size_t)STAGES * kStageBytes + 1024;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int tiles = 48, ctas = PAIR ? sms / 2 * 2 : sms;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(96);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, fn, mA, mB, a_blocks, 4, clk);     // warm-up
    cudaLaunchKernelEx(&cfg, fn, mA, mB, a_blocks, tiles, clk);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); exit(1); }
    long long* h = (long long*)malloc(ctas * sizeof(long long));
    cudaMemcpy(h, clk, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < ctas; ++i) avg += (double)h[i];
    avg /= ctas;
    const double per_stage = avg / (tiles * kKB);
    const double mma_clk = MODE == REL_MMA ? 4.0 * (BN / 2 > 64 ? BN / 2 : 64) : 0.0;
    printf("%-58s %2d x %2d KB  %7.0f clk/stage  %6.1f B/clk/SM", name, STAGES, kStageBytes / 1024, per_stage, kStageBytes / per_stage);
    if (MODE == REL_MMA) printf("  tensor pipe %4.0f %%", 100.0 * mma_clk / per_stage);
    printf("\n");
    free(h);
}

template <bool EPIW, bool EPIWORK, bool WDOG, bool RT, bool PROD2>
void run2(const char* name, EncodeTiledFn enc, void* pA, void* pB, long long* clk, int sms, int a_blocks) {
    CUtensorMap mA, mB;
    if (make_map(enc, &mA, pA, (long long)a_blocks * 2 * 128, 128) || make_map(enc, &mB, pB, 256, 128)) { printf("encode failed\n"); return; }
    auto fn = k2<EPIW, EPIWORK, WDOG, RT, PROD2>;
    const RtArgs rt{5, 32768, 16};
    const size_t smem = 5 * 32768 + 12 * 4096 + 1024;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int tiles = 48, ctas = sms / 2 * 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, fn, mA, mB, a_blocks, 4, clk, rt);     // warm-up
    cudaLaunchKernelEx(&cfg, fn, mA, mB, a_blocks, tiles, clk, rt);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); exit(1); }
    long long* h = (long long*)malloc(ctas * sizeof(long long));
    cudaMemcpy(h, clk, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < ctas; ++i) avg += (double)h[i];
    avg /= ctas;
    const double per_stage = avg / (tiles * 16);
    printf("%-58s  5 x 32 KB  %7.0f clk/stage  %6.1f B/clk/SM  tensor pipe %4.0f %%\n", name, per_stage, 32768 / per_stage, 100.0 * 512 / per_stage);
    free(h);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int a_blocks = 64 * 2;                  // 32 MB of A: L2 resident
    void *pA, *pB;
    long long* clk;
    cudaMalloc(&pA, (size_t)a_blocks * 128 * 2048);
    cudaMalloc(&pB, 256 * 2048);
    cudaMalloc(&clk, 256 * sizeof(long long));
    cudaMemset(pA, 0, (size_t)a_blocks * 128 * 2048);
    cudaMemset(pB, 0, 256 * 2048);
    printf("%d SMs, operands L2 resident, K = 1024 per tile\n", sms);
    run<false, REL_SOFT, 128, 6>("single CTA, software release", enc, pA, pB, clk, sms, a_blocks / 2);
    run<false, REL_COMMIT, 128, 6>("single CTA, tcgen05.commit release, no MMA", enc, pA, pB, clk, sms, a_blocks / 2);
    run<false, REL_MMA, 128, 6>("single CTA, 4 x mma 128x128x16 + commit", enc, pA, pB, clk, sms, a_blocks / 2);
    run<false, REL_MMA, 256, 4>("single CTA, 4 x mma 128x256x16 + commit", enc, pA, pB, clk, sms, a_blocks / 2);
    run<true, REL_COMMIT, 256, 5>("CTA pair, 2SM loads, commit multicast release, no MMA", enc, pA, pB, clk, sms, a_blocks / 2);
    run<true, REL_MMA, 256, 5>("CTA pair, 4 x mma.cta_group::2 256x256x16 + commit", enc, pA, pB, clk, sms, a_blocks / 2);
    run<true, REL_MMA, 256, 6>("CTA pair, same with 6 stages", enc, pA, pB, clk, sms, a_blocks / 2);
    run<true, REL_MMA, 256, 3>("CTA pair, same with 3 stages", enc, pA, pB, clk, sms, a_blocks / 2);
    if (getenv("RING_V2")) {   // engine look-alikes (see k2): one feature at a time, then all of them
        run2<false, false, false, false, false>("pair v2: 16-warp CTA, nothing else", enc, pA, pB, clk, sms, a_blocks / 2);
        run2<false, false, true, false, false>("pair v2: + watchdog waits", enc, pA, pB, clk, sms, a_blocks / 2);
        run2<false, false, false, true, false>("pair v2: + runtime stage geometry", enc, pA, pB, clk, sms, a_blocks / 2);
        run2<false, false, false, false, true>("pair v2: + second producer warp", enc, pA, pB, clk, sms, a_blocks / 2);
        run2<true, false, false, false, false>("pair v2: + 12 polling epilogue warps, 2 accumulators", enc, pA, pB, clk, sms, a_blocks / 2);
        run2<true, true, false, false, false>("pair v2: + epilogue work (ld, tanh, pack, sts)", enc, pA, pB, clk, sms, a_blocks / 2);
        run2<true, true, true, true, true>("pair v2: all of the above", enc, pA, pB, clk, sms, a_blocks / 2);
    }
    return 0;
}
