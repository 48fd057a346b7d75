// Micro-benchmark (B200): how many operand bytes per clock can TMA land in one SM's shared memory, and is the limit
// per SM or chip-wide?  The GEMM engine (csrc/gemm_tc.cu) measures 42 B/clk/SM with all 148 SMs loading (11.7 TB/s,
// profiles/r1s3_gemm_experiments.txt), which caps a 128 x 256 tile per CTA at ~2/3 of the tensor peak.  This probe runs
// the engine's load pipeline alone (ring of 32 KB stages, one elected producer lane, a consumer warp that only recycles
// the stages) and varies
//   * the number of CTAs (one per SM): 148 / 74 / 36 / 16     -> per-SM limit (rate stays) or chip limit (rate rises)?
//   * where the A operand lives: 16 MB (L2 resident) or 512 MB (streamed from HBM); B is always a 512 KB L2-resident tile
//   * unicast (each CTA fetches A 16 KB + B 16 KB per stage) against 2-CTA clusters in which each CTA fetches half of B
//     and multicasts it to both (24 KB requested, 32 KB landed per CTA and stage)
//   * 4-CTA clusters sharing B the same way (20 KB requested per CTA and stage; only 132 SMs can hold 4-clusters)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/l2_ingest tools/ubench/l2_ingest.cu -lcuda
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
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(bar), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        if (++spins > (1u << 26)) __trap();      // a pipeline bug becomes a CUDA error, not a hung GPU
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}

constexpr int kStages = 6, kStageBytes = 32768, kKB = 16;   // K = 1024 = 16 k-blocks of 64 bf16

// CL = cluster size (1, 2, 4).  Every CTA lands A [128 x 64] (16 KB, its own rows) + B [128 x 64] (16 KB, shared by the
// cluster when CL > 1: CTA r fetches rows [r * 128 / CL, (r + 1) * 128 / CL) and multicasts them).
template <int CL>
__global__ void __launch_bounds__(64, 1)
k(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB, int a_blocks, int tiles, long long* clk) {
    extern __shared__ uint8_t dsm[];
    __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
    const uint32_t base = (smem_u32(dsm) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), CL);      // every CTA of the cluster must have consumed the shared B stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (CL > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();
    const long long t0 = clock64();
    uint32_t stage = 0, phase = 0;
    if (warp == 0) {
        uint32_t el;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
        for (int t = 0; t < tiles; ++t) {
            const int blk = (int)((blockIdx.x + (long long)t * gridDim.x) % a_blocks);
            for (int kb = 0; kb < kKB; ++kb) {
                mbar_wait(smem_u32(&empty[stage]), phase ^ 1u);
                if (el) {
                    const uint32_t bar = smem_u32(&full[stage]), dst = base + stage * kStageBytes;
                    mbar_expect(bar, kStageBytes);
                    tma_load_2d(dst, &mA, bar, kb * 64, blk * 128);
                    if (CL == 1) tma_load_2d(dst + 16384, &mB, bar, kb * 64, 0);
                    else tma_load_2d_mc(dst + 16384 + rank * (16384 / CL), &mB, bar, kb * 64, rank * (128 / CL), (uint16_t)((1u << CL) - 1u));
                }
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        for (int t = 0; t < tiles; ++t)
            for (int kb = 0; kb < kKB; ++kb) {
                mbar_wait(smem_u32(&full[stage]), phase);
                if (lane == 0) {
                    if (CL == 1) mbar_arrive(smem_u32(&empty[stage]));
                    else for (uint32_t r = 0; r < CL; ++r) mbar_arrive_remote(smem_u32(&empty[stage]), r);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
    }
    __syncthreads();
    if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
    if (CL > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
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

template <int CL>
void run(EncodeTiledFn enc, void* pA, void* pB, long long* clk, int ctas, int a_blocks, const char* where) {
    CUtensorMap mA, mB;
    if (make_map(enc, &mA, pA, (long long)a_blocks * 128, 128) || make_map(enc, &mB, pB, 128, 128 / CL)) { printf("encode failed\n"); return; }
    auto fn = k<CL>;
    const size_t smem = kStages * kStageBytes + 1024;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (CL > 4) cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int tiles = 48;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(ctas / CL * CL));
    cfg.blockDim = dim3(64);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchKernelEx(&cfg, fn, mA, mB, a_blocks, 8, clk);     // warm-up
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, fn, mA, mB, a_blocks, tiles, clk);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("cluster %d ctas %d: %s\n", CL, ctas, cudaGetErrorString(err)); exit(1); }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const int n = ctas / CL * CL;
    long long* h = (long long*)malloc(n * sizeof(long long));
    cudaMemcpy(h, clk, n * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < n; ++i) avg += (double)h[i];
    avg /= n;
    const double landed = (double)tiles * kKB * kStageBytes;                    // bytes landed per CTA
    const double requested = (double)tiles * kKB * (16384.0 + 16384.0 / CL);    // bytes this CTA asked L2 for
    printf("cluster %d  %3d CTAs  A %-14s landed %6.1f B/clk/SM (%7.1f GB/s chip)   requested %6.1f B/clk/SM\n", CL, n, where,
           landed / avg, landed * n / (ms * 1e6), requested / avg);
    free(h);
}

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int big_blocks = 2048, small_blocks = 64;        // 512 MB / 16 MB of A
    void *pA, *pB;
    long long* clk;
    cudaMalloc(&pA, (size_t)big_blocks * 128 * 2048);
    cudaMalloc(&pB, 128 * 2048);
    cudaMalloc(&clk, 256 * sizeof(long long));
    cudaMemset(pA, 0, (size_t)big_blocks * 128 * 2048);
    cudaMemset(pB, 0, 128 * 2048);
    for (int pass = 0; pass < 2; ++pass) {
        const int blocks = pass == 0 ? small_blocks : big_blocks;
        const char* where = pass == 0 ? "16 MB (L2)" : "512 MB (HBM)";
        for (int ctas : {148, 74, 36, 16}) run<1>(enc, pA, pB, clk, ctas, blocks, where);
        for (int ctas : {148, 74, 36, 16}) run<2>(enc, pA, pB, clk, ctas, blocks, where);
        for (int ctas : {132, 64, 16}) run<4>(enc, pA, pB, clk, ctas, blocks, where);
    }
    return 0;
}
