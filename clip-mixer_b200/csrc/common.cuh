// Shared device/host helpers for libmixerclip (sm_100a only).
//
// PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / st)
// plus small math helpers (QuickGELU of training/clip/model.py:175-177 and its derivative).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mixerclip.h"

namespace mc {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing (no exceptions cross the C ABI; see include/mixerclip.h)
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define MC_CHECK(cond, ...)                                   \
    do {                                                      \
        if (!(cond)) {                                        \
            ::mc::set_last_error(__VA_ARGS__);                \
            return MC_ERR_INVALID;                            \
        }                                                     \
    } while (0)

#define MC_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::mc::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                                 __FILE__, __LINE__);                                        \
            return MC_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

int sm_count();

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the library executes MC_PDL_PROLOGUE() (griddepcontrol.wait)
// before its first global-memory access, and the hot kernels (GEMM engine, fused token mixing, LayerNorm) execute
// pdl_launch_dependents() when a CTA has finished its main loop.  With MC_PDL=1 every launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization: kernel N+1's grid is then launched and its CTAs scheduled (running
// their barrier / TMEM set-up) while kernel N drains its last stores, instead of paying the full launch latency at every
// one of the ~360 kernel boundaries of a step.  griddepcontrol.wait returns only when the preceding grid has completed
// and its memory operations are visible, so the ordering is that of plain stream serialisation; both instructions are
// no-ops in a kernel launched without the attribute.  (The trigger sits at the END of the main loops on purpose: fired
// at kernel entry, early CTAs of the dependent grid would squat on the SMs of the OTHER tower's stream.)
// ---------------------------------------------------------------------------------------------
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define MC_PDL_PROLOGUE() ::mc::pdl_wait()

template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}
#define MC_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (void)::mc::launch_k(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)
#endif

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
constexpr float kGeluA = 1.702f;  // model.py:177
constexpr float kLnEps = 1e-5f;   // nn.LayerNorm default (model.py:166)

// precise variants (fp32 validation mode and all row-wise kernels)
__device__ __forceinline__ float sigmoid_precise(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float quick_gelu_precise(float z) { return z * sigmoid_precise(kGeluA * z); }
__device__ __forceinline__ float quick_gelu_grad_precise(float z) {
    float s = sigmoid_precise(kGeluA * z);
    return s * (1.0f + kGeluA * z * (1.0f - s));
}

// fast variants for the tensor-core epilogues: sigmoid(x) = 0.5*tanh(0.5x)+0.5, one MUFU op
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ float quick_gelu_fast(float z) { return z * sigmoid_fast(kGeluA * z); }
__device__ __forceinline__ float quick_gelu_grad_fast(float z) {
    float s = sigmoid_fast(kGeluA * z);
    return s * fmaf(kGeluA * z, 1.0f - s, 1.0f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// ---------------------------------------------------------------------------------------------
// PTX: shared-memory addresses, mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
// Wait with a watchdog: a pipeline bug becomes a trap (CUDA error) after ~4 s instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (globaltimer_ns() - t0 > 4000000000ull) {
            printf("[mixerclip] mbarrier timeout: block (%d,%d,%d) thread %d bar 0x%x parity %u\n", blockIdx.x,
                   blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
            __trap();
        }
    }
}

// Wait for warps that are not on the critical path of a pipeline (epilogue warps polling for an accumulator): backs
// off with nanosleep so that the polling does not take issue slots from the warp that feeds the tensor pipe (the
// scheduler favours high warp ids, the MMA issuer is a single thread).  Watchdog by iteration count.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t sleep_ns) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(sleep_ns);
        if (++spins > (1u << 24)) {
            printf("[mixerclip] mbarrier timeout (relaxed): block (%d,%d,%d) thread %d bar 0x%x parity %u\n", blockIdx.x,
                   blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
            __trap();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// PTX: TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// PTX: tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 operands, fp32 accumulate.
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (base_lane + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA instruction descriptor, kind::f16 with bf16 A/B and fp32 D (bit layout: PTX ISA "Instruction
// descriptor"; c_format[4,6)=1 F32, a_format[7,10)=1 BF16, b_format[10,13)=1 BF16, a_major bit15,
// b_major bit16 (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) |
           (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// Shared-memory matrix descriptor for SWIZZLE_128B tiles (start>>4 at [0,14), LBO>>4 at [16,30),
// SBO>>4 at [32,46), version=1 at [46,48), layout_type=2 (128B swizzle) at [61,64)).
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return uint64_t((smem_addr & 0x3ffffu) >> 4) | (uint64_t((lbo_bytes >> 4) & 0x3fffu) << 16) |
           (uint64_t((sbo_bytes >> 4) & 0x3fffu) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}

}  // namespace mc
