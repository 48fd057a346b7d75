// Micro-benchmark (B200): how fast can the epilogue warps of a GEMM CTA push a 128 x 256 output tile to global memory?
// 148 CTAs x 12 warps (three per 32-row quarter, like csrc/gemm_tc.cu); each warp owns [32 rows x W columns] chunks,
// writes them into a swizzled staging tile and hands the box to TMA (cp.async.bulk.tensor store), or stores its row
// segment directly with 256-bit st.global.  Variants: element size, box width (64-byte vs 128-byte rows), one or two
// output tensors per chunk (C + saved pre-activation), one or two staging buffers per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tma_store_rate tools/ubench/tma_store_rate.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

constexpr int kWarps = 12, kTileRows = 128, kTileCols = 256;

// mode 0: TMA store; mode 1: direct 256-bit global stores
template <int ESZ, int W, int NT, int NBUF, int MODE>
__global__ void __launch_bounds__(kWarps * 32, 1)
k(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1, uint8_t* p0, uint8_t* p1, int tiles_per_cta,
  int iters, long long* clk) {
    extern __shared__ uint8_t dsm[];
    const uint32_t base = (smem_u32(dsm) + 1023u) & ~1023u;
    constexpr int kRowBytes = W * ESZ;                 // 64 or 128
    constexpr int kBoxBytes = 32 * kRowBytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, h = warp >> 2;
    const uint32_t stage0 = base + warp * (NBUF * NT * kBoxBytes);
    const uint32_t sw = kRowBytes == 128 ? (lane & 7) : ((lane >> 1) & 3);
    const long long ld = (long long)kTileCols * ESZ;
    long long t0 = clock64();
    int buf = 0;
    for (int it = 0; it < iters; ++it) {
        const int tile = blockIdx.x * tiles_per_cta + (it % tiles_per_cta);
        const int row0 = tile * kTileRows + q * 32;
        for (int c = h * W; c < kTileCols; c += 3 * W) {
            uint32_t v[kRowBytes / 4];
#pragma unroll
            for (int i = 0; i < kRowBytes / 4; ++i) v[i] = it * 977u + c + i + lane;
            if (MODE == 0) {
                if (lane == 0) {
                    if (NBUF == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                }
                __syncwarp();
                const uint32_t st = stage0 + buf * (NT * kBoxBytes);
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const uint32_t rowp = st + t * kBoxBytes + lane * kRowBytes;
#pragma unroll
                    for (int j = 0; j < kRowBytes / 16; ++j)
                        sts128(rowp + ((j ^ sw) << 4), v[4 * j] + t, v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&m0, st, c, row0, 0);
                    if (NT == 2) tma_store_3d(&m1, st + kBoxBytes, c, row0, 0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                buf = (buf + 1) % NBUF;
            } else {
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    uint8_t* gp = (t == 0 ? p0 : p1) + (long long)(row0 + lane) * ld + (long long)c * ESZ;
#pragma unroll
                    for (int j = 0; j < kRowBytes / 32; ++j)
                        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(gp + 32 * j), "r"(v[8 * j] + t),
                                     "r"(v[8 * j + 1]), "r"(v[8 * j + 2]), "r"(v[8 * j + 3]), "r"(v[8 * j + 4]), "r"(v[8 * j + 5]),
                                     "r"(v[8 * j + 6]), "r"(v[8 * j + 7]) : "memory");
                }
            }
        }
    }
    if (MODE == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ESZ, int W, int NT, int NBUF, int MODE>
void run(const char* name, EncodeTiledFn enc, uint8_t* p0, uint8_t* p1, long long* clk, int sms, int tiles_per_cta) {
    const long long rows = (long long)sms * tiles_per_cta * kTileRows;
    CUtensorMap m[2];
    for (int t = 0; t < 2; ++t) {
        cuuint64_t gdim[3] = {(cuuint64_t)kTileCols, (cuuint64_t)rows, 1};
        cuuint64_t gstride[2] = {(cuuint64_t)kTileCols * ESZ, (cuuint64_t)kTileCols * ESZ * rows};
        cuuint32_t box[3] = {W, 32, 1}, estr[3] = {1, 1, 1};
        CUresult r = enc(&m[t], ESZ == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, t == 0 ? p0 : p1,
                         gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         W * ESZ == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return; }
    }
    auto fn = k<ESZ, W, NT, NBUF, MODE>;
    const size_t smem = kWarps * NBUF * NT * 32 * W * ESZ + 1024;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 4 * tiles_per_cta;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fn<<<sms, kWarps * 32, smem>>>(m[0], m[1], p0, p1, tiles_per_cta, tiles_per_cta, clk);   // warm-up
    cudaEventRecord(e0);
    fn<<<sms, kWarps * 32, smem>>>(m[0], m[1], p0, p1, tiles_per_cta, iters, clk);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); exit(1); }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long* h = (long long*)malloc(sms * sizeof(long long));
    cudaMemcpy(h, clk, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    const double bytes_cta = (double)iters * kTileRows * kTileCols * ESZ * NT;
    printf("%-46s %8.0f clk/tile  %6.1f B/clk/SM  %7.1f GB/s\n", name, avg / iters, bytes_cta / avg, bytes_cta * sms / (ms * 1e6));
    free(h);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, tiles_per_cta = 16;
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const size_t bytes = (size_t)sms * tiles_per_cta * kTileRows * kTileCols * 4;   // 310 MB per tensor > L2
    uint8_t *p0, *p1;
    long long* clk;
    cudaMalloc(&p0, bytes); cudaMalloc(&p1, bytes); cudaMalloc(&clk, sms * sizeof(long long));
    printf("%d SMs, %d tiles of 128 x 256 per CTA, 12 warps per CTA\n", sms, tiles_per_cta);
    run<2, 32, 2, 1, 0>("TMA bf16 box 32x32 (64 B rows), C+Z, 1 buf", enc, p0, p1, clk, sms, tiles_per_cta);
    run<2, 32, 2, 2, 0>("TMA bf16 box 32x32 (64 B rows), C+Z, 2 buf", enc, p0, p1, clk, sms, tiles_per_cta);
    run<2, 64, 2, 1, 0>("TMA bf16 box 32x64 (128 B rows), C+Z, 1 buf", enc, p0, p1, clk, sms, tiles_per_cta);
    run<2, 64, 2, 2, 0>("TMA bf16 box 32x64 (128 B rows), C+Z, 2 buf", enc, p0, p1, clk, sms, tiles_per_cta);
    run<2, 32, 1, 1, 0>("TMA bf16 box 32x32 (64 B rows), C, 1 buf", enc, p0, p1, clk, sms, tiles_per_cta);
    run<2, 32, 1, 2, 0>("TMA bf16 box 32x32 (64 B rows), C, 2 buf", enc, p0, p1, clk, sms, tiles_per_cta);
    run<2, 64, 1, 1, 0>("TMA bf16 box 32x64 (128 B rows), C, 1 buf", enc, p0, p1, clk, sms, tiles_per_cta);
    run<2, 64, 1, 2, 0>("TMA bf16 box 32x64 (128 B rows), C, 2 buf", enc, p0, p1, clk, sms, tiles_per_cta);
    run<4, 32, 1, 1, 0>("TMA fp32 box 32x32 (128 B rows), 1 buf", enc, p0, p1, clk, sms, tiles_per_cta);
    run<4, 32, 1, 2, 0>("TMA fp32 box 32x32 (128 B rows), 2 buf", enc, p0, p1, clk, sms, tiles_per_cta);
    run<2, 32, 2, 1, 1>("st.global.v8 bf16, 64 B per lane, C+Z", enc, p0, p1, clk, sms, tiles_per_cta);
    run<2, 64, 2, 1, 1>("st.global.v8 bf16, 128 B per lane, C+Z", enc, p0, p1, clk, sms, tiles_per_cta);
    run<2, 32, 1, 1, 1>("st.global.v8 bf16, 64 B per lane, C", enc, p0, p1, clk, sms, tiles_per_cta);
    run<4, 32, 1, 1, 1>("st.global.v8 fp32, 128 B per lane", enc, p0, p1, clk, sms, tiles_per_cta);
    return 0;
}
