// Micro-benchmark (B200): clocks per tcgen05.mma (kind::f16, M = 128, K = 16, SS operands in 128B-swizzled smem)
// as a function of N, operand majors and of how many independent TMEM accumulators the instruction stream rotates
// over.  One CTA per SM, one issuing thread; garbage operands (timing only).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return uint64_t((addr & 0x3ffffu) >> 4) | (uint64_t((lbo >> 4) & 0x3fffu) << 16) | (uint64_t((sbo >> 4) & 0x3fffu) << 32) |
           (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) | (uint32_t(N >> 3) << 17) |
           (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}

// mode bits: a_mn = mode & 1, b_mn = (mode >> 1) & 1
__global__ void __launch_bounds__(384, 1) k(int N, int mode, int nacc, int iters, long long* out, int ldtm, int do_mma, int commit_every = 0) {
    extern __shared__ uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t bar2[8];
    __shared__ uint32_t tbase;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2[i])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + (base - smem_u32(smem)))[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tbase;
    if (threadIdx.x >= 128 && ldtm) {
        // 8 warps reading TMEM like the E1 warps of csrc/tokenmix.cu: 32 lanes x 32 columns per instruction
        const uint32_t q = (threadIdx.x >> 5) & 3;
        uint32_t acc = 0;
        long long t0 = clock64();
        for (int it = 0; it < ldtm; ++it) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(tm + ((q * 32) << 16) + 256 + (it & 3) * 32)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= r[i];
        }
        long long t1 = clock64();
        if (acc == 0x12345678u) out[3] = acc;
        if (blockIdx.x == 0 && threadIdx.x == 128) out[2] = t1 - t0;
    }
    if (threadIdx.x < 32 && do_mma) {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        const bool leader = pred != 0;
        const int a_mn = mode & 1, b_mn = (mode >> 1) & 1;
        const uint32_t id = idesc(128, N, a_mn, b_mn);
        const uint32_t a0 = base, b0 = base + 32 * 1024;
        // K-major: k-step +32 B, LBO 16, SBO 1024;  MN-major: k-step +2048 B, LBO = 8192 (64 k-rows x 128 B), SBO 1024
        const uint32_t akstep = a_mn ? 2048 : 32, bkstep = b_mn ? 2048 : 32;
        const uint32_t albo = a_mn ? 8192 : 16, blbo = b_mn ? 8192 : 16;
        const int acc_stride = N;
        const uint64_t ad = sdesc(a0, albo, 1024), bd = sdesc(b0, blbo, 1024);
        const uint32_t aks = akstep >> 4, bks = bkstep >> 4;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t d0 = tm + ((it % nacc) * acc_stride);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma(d0, ad + ks * aks, bd + ks * bks, id, 1u);
                if (commit_every && (it % commit_every) == 0)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[it & 7])) : "memory");
            }
        }
        if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main() {
    long long* d; long long h[4];
    cudaMalloc(&d, 32);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 256;
    printf("%-6s %-10s %-5s %12s %12s\n", "N", "majors", "nacc", "issue clk/mma", "total clk/mma");
    for (int N : {64, 128, 256})
        for (int mode : {0, 3})
            for (int nacc : {1, 2}) {
                if (nacc * N > 512) continue;
                k<<<148, 384, 200 * 1024>>>(N, mode, nacc, iters, d, 0, 1);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                const char* mj = mode == 0 ? "A:K B:K" : mode == 1 ? "A:MN B:K" : "A:MN B:MN";
                printf("%-6d %-10s %-5d %12.1f %12.1f   (floor %d)\n", N, mj, nacc, (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), N / 2);
            }
    // TMEM read / MMA interference
    printf("\nconcurrent tcgen05.ld (8 warps, 32x32b.x32 = 4 KB each) and tcgen05.mma (N = 64 / 128, A:K B:K)\n");
    for (int N : {64, 128}) {
        for (int cfg = 0; cfg < 3; ++cfg) {
            const int do_mma = cfg != 1, ldtm = cfg != 0 ? 4096 : 0;
            cudaMemset(d, 0, 32);
            k<<<148, 384, 200 * 1024>>>(N, 0, 2, 1024, d, ldtm, do_mma);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
            printf("N=%-4d mma=%d ldtm=%d : %8.1f clk/mma   %8.1f clk per tcgen05.ld.x32 per warp (8 warps)\n", N, do_mma, ldtm != 0,
                   do_mma ? (double)h[1] / 4096 : 0.0, ldtm ? (double)h[2] / ldtm : 0.0);
        }
    }
    printf("\ncommit cost: N = 64, 4 MMAs per group, tcgen05.commit every k-th group\n");
    for (int ce : {0, 4, 2, 1}) {
        k<<<148, 384, 200 * 1024>>>(64, 0, 2, 1024, d, 0, 1, ce);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("commit_every=%d : issue %8.1f clk/mma, total %8.1f clk/mma\n", ce, (double)h[0] / 4096, (double)h[1] / 4096);
    }
    return 0;
}
