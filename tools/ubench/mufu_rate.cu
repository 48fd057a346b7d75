// Micro-benchmark (B200): issue rate of the transcendental instructions used by the GELU epilogues and the cost of
// fence.proxy.async, measured per SM with 8 warps (2 per SMSP) like the E1 warps of csrc/tokenmix.cu.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
    __shared__ float sm[1024];
    sm[threadIdx.x] = x0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) {
            asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x0)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x1));
            asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x2)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x3));
        } else if (OP == 1) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
        } else if (OP == 2) {
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x0)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x2)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x3));
        } else if (OP == 3) {
            unsigned a = __float_as_uint(x0), b = __float_as_uint(x1), c = __float_as_uint(x2), d = __float_as_uint(x3);
            asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(b));
            asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(c)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(d));
            x0 = __uint_as_float(a); x1 = __uint_as_float(b); x2 = __uint_as_float(c); x3 = __uint_as_float(d);
        } else if (OP == 4) {
            unsigned a = __float_as_uint(x0), b = __float_as_uint(x1), c = __float_as_uint(x2), d = __float_as_uint(x3);
            asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b));
            asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(c)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(d));
            x0 = __uint_as_float(a); x1 = __uint_as_float(b); x2 = __uint_as_float(c); x3 = __uint_as_float(d);
        } else if (OP == 5) {
            sm[threadIdx.x] = x0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            x0 += 1.f;
        } else if (OP == 6) {
            x0 = fmaf(x0, x1, x2); x1 = fmaf(x1, x2, x3); x2 = fmaf(x2, x3, x0); x3 = fmaf(x3, x0, x1);
        } else if (OP == 7) {   // F2FP.BF16: 4 independent packs per iteration
            unsigned a, b, c, d;
            asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a) : "f"(x0), "f"(x1));
            asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b) : "f"(x1), "f"(x2));
            asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(c) : "f"(x2), "f"(x3));
            asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x3), "f"(x0));
            x0 = __uint_as_float(a ^ 0x3f000000u); x1 = __uint_as_float(b ^ 0x3f000000u);
            x2 = __uint_as_float(c ^ 0x3f000000u); x3 = __uint_as_float(d ^ 0x3f000000u);
        } else if (OP == 8) {   // F2FP.F16
            unsigned a, b, c, d;
            asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a) : "f"(x0), "f"(x1));
            asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(b) : "f"(x1), "f"(x2));
            asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(c) : "f"(x2), "f"(x3));
            asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x3), "f"(x0));
            x0 = __uint_as_float(a ^ 0x3f000000u); x1 = __uint_as_float(b ^ 0x3f000000u);
            x2 = __uint_as_float(c ^ 0x3f000000u); x3 = __uint_as_float(d ^ 0x3f000000u);
        } else if (OP == 9) {   // packed fp32 FMA
            float2 a = make_float2(x0, x1), b = make_float2(x2, x3);
            a = __ffma2_rn(a, b, a); b = __ffma2_rn(b, a, b); a = __ffma2_rn(a, b, b); b = __ffma2_rn(b, a, a);
            x0 = a.x; x1 = a.y; x2 = b.x; x3 = b.y;
        } else if (OP == 10) {  // f16 -> f32 unpack
            unsigned a = __float_as_uint(x0), b = __float_as_uint(x1);
            float f0, f1, f2, f3;
            asm volatile("{.reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(f0), "=f"(f1) : "r"(a));
            asm volatile("{.reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(f2), "=f"(f3) : "r"(b));
            x0 = f0 + x2; x1 = f1 + x3; x2 = f2; x3 = f3;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + sm[(threadIdx.x + 1) & 1023];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name, int per_iter) {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<OP><<<148, 256>>>(out, iters, cyc);
    k<OP><<<148, 256>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double lanes = 256.0 * iters * per_iter;
    printf("%-22s %8.1f cycles/iter/warp-pair-per-SMSP   %6.2f lane-ops/clk/SM\n", name, (double)h / iters, lanes / h);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("tanh.approx.f32", 4); run<1>("ex2.approx.ftz.f32", 4); run<2>("rcp.approx.ftz.f32", 4);
    run<3>("tanh.approx.f16x2", 4); run<4>("ex2.approx.f16x2", 4); run<5>("sts+fence.proxy.async", 1); run<6>("ffma", 4);
    run<7>("cvt.rn.bf16x2.f32", 4); run<8>("cvt.rn.f16x2.f32", 4); run<9>("ffma2 (dep chain)", 4); run<10>("cvt.f32.f16 x4", 4);
    return 0;
}
