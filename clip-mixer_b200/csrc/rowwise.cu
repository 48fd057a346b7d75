// Row-wise (HBM-bound) kernels of the Mixer-CLIP step: LayerNorm forward / backward with the
// reductions that ride on it, bias-gradient reductions, im2col, embedding gather / scatter-add,
// L2 normalisation, casts.  One warp per row, 16-byte accesses, fp32 math throughout.
// Reference call sites: training/clip/model.py:166-172 (LayerNorm), :258,272 (patch conv),
// :275-277 (class token), :414 (token embedding), :424 (EOT argmax), :433-434 (normalise).
#include "common.cuh"

namespace mc {
namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_act4(void* base, int dtype, long long idx, float4 v) {
    if (dtype == MC_BF16) {
        uint2 pk;
        pk.x = pack_bf16x2(v.x, v.y);
        pk.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = pk;
    } else {
        st4(reinterpret_cast<float*>(base) + idx, v);
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm forward
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ln_fwd_kernel(const float* __restrict__ x, long long x_row_stride, const int* __restrict__ row_index,
              const float* __restrict__ cls, long long cls_period, const float* __restrict__ gamma,
              const float* __restrict__ beta, void* __restrict__ y, int y_dtype, long long y_row_stride,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, int D) {
    MC_PDL_PROLOGUE();
    const long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = (cls != nullptr && row % cls_period == 0)
                          ? cls
                          : x + (row_index ? (long long)row_index[row] : row) * x_row_stride;
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        v[i] = c < D ? ld4(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    pdl_launch_dependents();      // this row's loads are in flight; the kernel is one wave or two of short CTAs
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < D) {
            const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
            q += a * a + b * b + cc * cc + d * d;
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / D + kLnEps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < D) {
            const float4 g = ld4(gamma + c), b = ld4(beta + c);
            float4 o;
            o.x = (v[i].x - mean) * rstd * g.x + b.x;
            o.y = (v[i].y - mean) * rstd * g.y + b.y;
            o.z = (v[i].z - mean) * rstd * g.z + b.z;
            o.w = (v[i].w - mean) * rstd * g.w + b.w;
            st_act4(y, y_dtype, row * y_row_stride + c, o);
        }
    }
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward (+ residual grad, dgamma/dbeta, column sums, row sums, class-token grad)
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, long long x_row_stride,
              const int* __restrict__ row_index, const float* __restrict__ cls, long long cls_period,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              const float* dres, float* dx, long long dx_row_stride,   // dres may alias dx (in-place residual add): no __restrict__
              void* __restrict__ dx_act, int act_dtype, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ colsum_out, float* __restrict__ rowsum_out, long long rowsum_period,
              float* __restrict__ dcls, long long rows, int D) {
    MC_PDL_PROLOGUE();
    __shared__ float red[kWarpsPerBlock][VPL * 128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 ag[VPL], ab[VPL], ae[VPL];  // dgamma, dbeta, extra (colsum or class-token grad)
    float4 gm[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        ag[i] = ab[i] = ae[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = (i * 32 + lane) * 4;
        gm[i] = c < D ? ld4(gamma + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const bool want_extra = (dcls != nullptr) || (colsum_out != nullptr);
    const float invD = 1.0f / D;
    for (long long row = (long long)blockIdx.x * kWarpsPerBlock + warp; row < rows;
         row += (long long)gridDim.x * kWarpsPerBlock) {
        const bool is_cls = (cls != nullptr && row % cls_period == 0);
        const long long pos = row_index ? (long long)row_index[row] : row;
        const float* xr = is_cls ? cls : x + pos * x_row_stride;
        const float* dyr = dy + row * (long long)D;
        const float mu = mean[row], rs = rstd[row];
        float4 xh[VPL], d[VPL];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                const float4 xv = ld4(xr + c), g = ld4(dyr + c);
                xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                ag[i].x += g.x * xh[i].x; ag[i].y += g.y * xh[i].y; ag[i].z += g.z * xh[i].z; ag[i].w += g.w * xh[i].w;
                ab[i].x += g.x; ab[i].y += g.y; ab[i].z += g.z; ab[i].w += g.w;
                d[i] = make_float4(g.x * gm[i].x, g.y * gm[i].y, g.z * gm[i].z, g.w * gm[i].w);
                s1 += d[i].x + d[i].y + d[i].z + d[i].w;
                s2 += d[i].x * xh[i].x + d[i].y * xh[i].y + d[i].z * xh[i].z + d[i].w * xh[i].w;
            } else {
                xh[i] = d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        const float c1 = warp_sum(s1) * invD, c2 = warp_sum(s2) * invD;
        float rsum = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                float4 o;
                o.x = rs * (d[i].x - c1 - xh[i].x * c2);
                o.y = rs * (d[i].y - c1 - xh[i].y * c2);
                o.z = rs * (d[i].z - c1 - xh[i].z * c2);
                o.w = rs * (d[i].w - c1 - xh[i].w * c2);
                if (dres != nullptr) {
                    const float4 r = ld4(dres + pos * dx_row_stride + c);
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                if (is_cls) {
                    ae[i].x += o.x; ae[i].y += o.y; ae[i].z += o.z; ae[i].w += o.w;
                    // class-token rows are parameters, not activations: their operand copy must be zero
                    // so that the patch-embedding weight gradient does not see them
                    if (dx_act != nullptr) st_act4(dx_act, act_dtype, pos * dx_row_stride + c, make_float4(0.f, 0.f, 0.f, 0.f));
                } else {
                    if (dx != nullptr) st4(dx + pos * dx_row_stride + c, o);
                    if (dx_act != nullptr) st_act4(dx_act, act_dtype, pos * dx_row_stride + c, o);
                    if (dcls == nullptr && colsum_out != nullptr) {
                        ae[i].x += o.x; ae[i].y += o.y; ae[i].z += o.z; ae[i].w += o.w;
                    }
                }
                rsum += o.x + o.y + o.z + o.w;
            }
        }
        if (rowsum_out != nullptr) {
            rsum = warp_sum(rsum);
            if (lane == 0) atomicAdd(rowsum_out + row % rowsum_period, rsum);
        }
    }
    // block reduction of the three accumulators, then one atomic per column per block
    for (int which = 0; which < 3; ++which) {
        float* out = which == 0 ? dgamma : which == 1 ? dbeta : (dcls != nullptr ? dcls : colsum_out);
        if (out == nullptr || (which == 2 && !want_extra)) continue;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float4 a = which == 0 ? ag[i] : which == 1 ? ab[i] : ae[i];
            st4(&red[warp][(i * 32 + lane) * 4], a);
        }
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += blockDim.x) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kWarpsPerBlock; ++w) s += red[w][c];
            atomicAdd(out + c, s);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward, dense case (the 48 block-level calls of a step): dx = dres + LNbwd(dy) (+ act copy, + row
// sums) with the column accumulators (dgamma, dbeta, column sums of dx) kept in shared memory, so the operands
// are read from HBM exactly once and the register footprint stays that of a plain row kernel.
// The fused kernel above keeps serving the gather / class-token variants (3 calls per step).
// ------------------------------------------------------------------------------------------------
constexpr int kRowsumLocal = 256;

template <int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 2)
ln_bwd_rows_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ gamma, const float* dres,
                   float* dx /* dres may alias dx: every element is read, then written, by the same thread */, void* __restrict__ dx_act, int act_dtype, float* __restrict__ rowsum_out,
                   long long rowsum_period, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   float* __restrict__ colsum_out, long long rows, int D) {
    MC_PDL_PROLOGUE();
    // Persistent over rows.  The column accumulators (dgamma, dbeta, column sums of dx) live in shared memory,
    // one private copy per warp (lane-owned columns: conflict-free read-modify-write), so the kernel keeps the
    // register footprint of a plain row kernel (full occupancy) and the operands are read from HBM exactly once.
    extern __shared__ float acc_sm[];                       // [warp][3][VPL*128]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int W = VPL * 128;
    float* my = acc_sm + (size_t)warp * 3 * W;
    const bool want_cs = colsum_out != nullptr;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        st4(my + c, make_float4(0.f, 0.f, 0.f, 0.f));
        st4(my + W + c, make_float4(0.f, 0.f, 0.f, 0.f));
        st4(my + 2 * W + c, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    // per-period row sums (the token-mixing lin2 bias gradient: 77 or 50 addresses for ~20 k rows): collected per block in
    // shared memory, one global atomic per address and block - the 19712 same-address global atomics of the text tower
    // cost 4 us per launch (profiles/r1s2_rowwise_bench.txt: +rowsum 39.6 us against +colsum 35.6 us)
    __shared__ float rs_sm[kRowsumLocal];
    const bool rs_local = rowsum_out != nullptr && rowsum_period <= kRowsumLocal;
    if (rs_local) {
        for (int i = threadIdx.x; i < (int)rowsum_period; i += blockDim.x) rs_sm[i] = 0.f;
        __syncthreads();
    }
    const float invD = 1.0f / D;
    for (long long row = (long long)blockIdx.x * kWarpsPerBlock + warp; row < rows;
         row += (long long)gridDim.x * kWarpsPerBlock) {
        const float* xr = x + row * (long long)D;
        const float* dyr = dy + row * (long long)D;
        const float* rr = dres + row * (long long)D;
        // every HBM operand of the row is requested before anything is consumed (18 x 16-byte loads in flight per
        // lane): with the loads interleaved chunk by chunk the kernel paid one DRAM round trip per chunk and sat at
        // 3.0 TB/s (ncu: all stalls long-scoreboard)
        float4 xh[VPL], d[VPL], r[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                xh[i] = ld4(xr + c);
                d[i] = ld4(dyr + c);
            }
        }
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = (i * 32 + lane) * 4;
            r[i] = (dres != nullptr && c < D) ? ld4(rr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float mu = mean[row], rs = rstd[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                const float4 xv = xh[i], g = d[i], gm = ld4(gamma + c);
                xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                float4 a = ld4(my + c), bsum = ld4(my + W + c);
                a.x += g.x * xh[i].x; a.y += g.y * xh[i].y; a.z += g.z * xh[i].z; a.w += g.w * xh[i].w;
                bsum.x += g.x; bsum.y += g.y; bsum.z += g.z; bsum.w += g.w;
                st4(my + c, a);
                st4(my + W + c, bsum);
                d[i] = make_float4(g.x * gm.x, g.y * gm.y, g.z * gm.z, g.w * gm.w);
                s1 += d[i].x + d[i].y + d[i].z + d[i].w;
                s2 += d[i].x * xh[i].x + d[i].y * xh[i].y + d[i].z * xh[i].z + d[i].w * xh[i].w;
            } else {
                xh[i] = d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        const float c1 = warp_sum(s1) * invD, c2 = warp_sum(s2) * invD;
        float rsum = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                float4 o;
                o.x = rs * (d[i].x - c1 - xh[i].x * c2) + r[i].x;
                o.y = rs * (d[i].y - c1 - xh[i].y * c2) + r[i].y;
                o.z = rs * (d[i].z - c1 - xh[i].z * c2) + r[i].z;
                o.w = rs * (d[i].w - c1 - xh[i].w * c2) + r[i].w;
                st4(dx + row * (long long)D + c, o);
                if (dx_act != nullptr) st_act4(dx_act, act_dtype, row * (long long)D + c, o);
                if (want_cs) {
                    float4 cs = ld4(my + 2 * W + c);
                    cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w;
                    st4(my + 2 * W + c, cs);
                }
                rsum += o.x + o.y + o.z + o.w;
            }
        }
        if (rowsum_out != nullptr) {
            rsum = warp_sum(rsum);
            if (lane == 0) {
                if (rs_local) atomicAdd(&rs_sm[(int)(row % rowsum_period)], rsum);
                else atomicAdd(rowsum_out + row % rowsum_period, rsum);
            }
        }
    }
    if (threadIdx.x == 0) pdl_launch_dependents();          // rows done; only the column reductions remain
    __syncthreads();
    if (rs_local)
        for (int i = threadIdx.x; i < (int)rowsum_period; i += blockDim.x)
            if (rs_sm[i] != 0.f) atomicAdd(rowsum_out + i, rs_sm[i]);
    for (int idx = threadIdx.x; idx < 3 * D; idx += blockDim.x) {
        const int which = idx / D, c = idx - which * D;
        float* out = which == 0 ? dgamma : which == 1 ? dbeta : colsum_out;
        if (out == nullptr) continue;
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) sum += acc_sm[(size_t)w * 3 * W + which * W + c];
        atomicAdd(out + c, sum);
    }
}

// ------------------------------------------------------------------------------------------------
// bias-gradient reductions
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// 8 consecutive values of an act tensor as floats (16-byte load for bf16, 2 x 16-byte for fp32)
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const float4 a = ld4(p), b = ld4(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    v[0] = bf16_lo(a.x); v[1] = bf16_hi(a.x); v[2] = bf16_lo(a.y); v[3] = bf16_hi(a.y);
    v[4] = bf16_lo(a.z); v[5] = bf16_hi(a.z); v[6] = bf16_lo(a.w); v[7] = bf16_hi(a.w);
}

// out[c] += sum_r x[r,c]: block = 32 column groups of 8 x 8 row lanes; grid (ceil(cols/256), row chunks).
// Requires cols % 8 == 0 and 16-byte aligned rows (host checks, else the scalar kernel below runs).
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ x, long long rows, long long cols, long long ld, float* __restrict__ out,
                  long long rows_per_block) {
    MC_PDL_PROLOGUE();
    __shared__ float red[8][32][9];
    const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const long long c = ((long long)blockIdx.x * 32 + cg) * 8;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < cols) {
#pragma unroll 4
        for (long long r = r0 + rl; r < r1; r += 8) {
            float v[8];
            load8<T>(x + r * ld + c, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += v[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[rl][cg][i] = acc[i];
    __syncthreads();
    // 256 threads -> 256 columns of the block
    const int col = threadIdx.x;
    const long long gc = (long long)blockIdx.x * 256 + col;
    if (gc < cols) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += red[k][col >> 3][col & 7];
        atomicAdd(out + gc, sum);
    }
}

template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long long rows, long long cols, long long ld, float* __restrict__ out,
                              long long rows_per_block) {
    MC_PDL_PROLOGUE();
    const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (c >= cols) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float s0 = 0.f, s1 = 0.f;
    const bool two = c + 1 < cols;
    for (long long r = r0; r < r1; ++r) {
        const T* p = x + r * ld + c;
        s0 += to_f<T>(p[0]);
        if (two) s1 += to_f<T>(p[1]);
    }
    atomicAdd(out + c, s0);
    if (two) atomicAdd(out + c + 1, s1);
}

// out[r % period] += sum_c x[r,c]; one warp per row, 16-byte loads when the row allows it
template <typename T>
__global__ void rowsum_kernel(const T* __restrict__ x, long long rows, long long cols, long long ld, long long period,
                              float* __restrict__ out, int vec) {
    MC_PDL_PROLOGUE();
    const long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const T* p = x + row * ld;
    float s = 0.f;
    if (vec) {
        for (long long c = lane * 8; c < cols; c += 256) {
            float v[8];
            load8<T>(p + c, v);
            s += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        }
    } else {
        for (long long c = lane; c < cols; c += 32) s += to_f<T>(p[c]);
    }
    s = warp_sum(s);
    if (lane == 0) atomicAdd(out + row % period, s);
}

// ------------------------------------------------------------------------------------------------
// casts / layout
// ------------------------------------------------------------------------------------------------
__global__ void cast_pad_kernel(const float* __restrict__ src, long long rows, long long cols, long long src_ld,
                                void* __restrict__ dst, int dst_dtype, long long dst_ld) {
    MC_PDL_PROLOGUE();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * dst_ld) return;
    const long long r = i / dst_ld, c = i % dst_ld;
    const float v = c < cols ? src[r * src_ld + c] : 0.f;
    if (dst_dtype == MC_BF16) reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(dst)[i] = v;
}

__constant__ float kImgMean[3] = {0.48145466f, 0.4578275f, 0.40821073f};  // training.py:115
__constant__ float kImgStd[3] = {0.26862954f, 0.26130258f, 0.27577711f};

// one thread = 4 consecutive pixels of a patch row (patch % 4 == 0): 4-byte (uint8) or 16-byte (fp32) load,
// 8-byte (bf16) or 16-byte (fp32) store, both coalesced
template <bool U8>
__global__ void im2col_kernel(const void* __restrict__ image, long long B, int R, int patch, void* __restrict__ out,
                              int out_dtype) {
    MC_PDL_PROLOGUE();
    const int g = R / patch;
    const long long Kc = 3ll * patch * patch;
    const long long total4 = B * g * g * Kc / 4;
    const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= total4) return;
    const long long i = i4 * 4;
    const long long row = i / Kc;
    const int col = (int)(i % Kc);
    const int px = col % patch, py = (col / patch) % patch, c = col / (patch * patch);
    const int gx = (int)(row % g), gy = (int)((row / g) % g);
    const long long b = row / (g * g);
    const long long src = ((b * 3 + c) * R + (gy * patch + py)) * (long long)R + gx * patch + px;
    float4 v;
    if (U8) {
        const uchar4 u = *reinterpret_cast<const uchar4*>(reinterpret_cast<const uint8_t*>(image) + src);
        const float m = kImgMean[c], is = 1.0f / kImgStd[c];
        v = make_float4((u.x * (1.0f / 255.0f) - m) * is, (u.y * (1.0f / 255.0f) - m) * is,
                        (u.z * (1.0f / 255.0f) - m) * is, (u.w * (1.0f / 255.0f) - m) * is);
    } else {
        v = ld4(reinterpret_cast<const float*>(image) + src);
    }
    st_act4(out, out_dtype, i, v);
}

// ------------------------------------------------------------------------------------------------
// token embedding
// ------------------------------------------------------------------------------------------------
__global__ void embed_fwd_kernel(const long long* __restrict__ text, const float* __restrict__ table, float* __restrict__ x,
                                 long long rows, int W, long long vocab) {
    MC_PDL_PROLOGUE();
    const long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    long long tok = text[row];
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
    const float* src = table + tok * W;
    float* dst = x + row * W;
    for (int c = lane * 4; c < W; c += 128) st4(dst + c, ld4(src + c));
}

// one warp per sample; consecutive positions holding the same token (padding after EOT) are summed in
// registers before a single vector atomic flush
template <int VPL>
__global__ void embed_bwd_kernel(const long long* __restrict__ text, const float* __restrict__ dx, float* __restrict__ dtable,
                                 long long B, int C, int W, long long vocab) {
    MC_PDL_PROLOGUE();
    const long long b = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    float4 acc[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    long long cur = -1;
    for (int t = 0; t <= C; ++t) {
        long long tok = -2;
        if (t < C) {
            tok = text[b * C + t];
            tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
        }
        if (tok != cur) {
            if (cur >= 0) {
#pragma unroll
                for (int i = 0; i < VPL; ++i) {
                    const int c = (i * 32 + lane) * 4;
                    if (c < W) atomicAdd(reinterpret_cast<float4*>(dtable + cur * W + c), acc[i]);
                    acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            cur = tok;
        }
        if (t < C) {
            const float* src = dx + (b * C + t) * (long long)W;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int c = (i * 32 + lane) * 4;
                if (c < W) {
                    const float4 v = ld4(src + c);
                    acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
                }
            }
        }
    }
}

__global__ void eot_rows_kernel(const long long* __restrict__ text, int* __restrict__ eot_row, long long B, int C) {
    MC_PDL_PROLOGUE();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    long long best = text[b * C];
    int arg = 0;
    for (int t = 1; t < C; ++t) {
        const long long v = text[b * C + t];
        if (v > best) { best = v; arg = t; }
    }
    eot_row[b] = (int)(b * C + arg);
}

// ------------------------------------------------------------------------------------------------
// L2 normalisation
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void l2norm_fwd_kernel(const float* __restrict__ f, float* __restrict__ u, float* __restrict__ inv_norm,
                                  long long rows, int E) {
    MC_PDL_PROLOGUE();
    const long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        v[i] = c < E ? ld4(f + row * E + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        s += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    const float inv = 1.0f / sqrtf(warp_sum(s));
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < E) st4(u + row * E + c, make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv));
    }
    if (lane == 0) inv_norm[row] = inv;
}

template <int VPL>
__global__ void l2norm_bwd_kernel(const float* __restrict__ du, const float* __restrict__ u, const float* __restrict__ inv_norm,
                                  float* __restrict__ df, void* __restrict__ df_act, int act_dtype, long long rows, int E) {
    MC_PDL_PROLOGUE();
    const long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float4 g[VPL], uv[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < E) {
            g[i] = ld4(du + row * E + c);
            uv[i] = ld4(u + row * E + c);
            s += g[i].x * uv[i].x + g[i].y * uv[i].y + g[i].z * uv[i].z + g[i].w * uv[i].w;
        } else {
            g[i] = uv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const float dot = warp_sum(s), inv = inv_norm[row];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < E) {
            const float4 o = make_float4((g[i].x - uv[i].x * dot) * inv, (g[i].y - uv[i].y * dot) * inv,
                                         (g[i].z - uv[i].z * dot) * inv, (g[i].w - uv[i].w * dot) * inv);
            if (df != nullptr) st4(df + row * E + c, o);
            if (df_act != nullptr) st_act4(df_act, act_dtype, row * E + c, o);
        }
    }
}

inline int pick_vpl(int64_t D) {
    const int64_t need = ceil_div(D, 128);
    if (need <= 1) return 1;
    if (need <= 2) return 2;
    if (need <= 4) return 4;
    if (need <= 6) return 6;
    if (need <= 8) return 8;
    return 0;
}

#define MC_DISPATCH_VPL(vpl, ...)                     \
    switch (vpl) {                                    \
        case 1: { constexpr int VPL = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int VPL = 2; __VA_ARGS__; } break; \
        case 4: { constexpr int VPL = 4; __VA_ARGS__; } break; \
        case 6: { constexpr int VPL = 6; __VA_ARGS__; } break; \
        default: { constexpr int VPL = 8; __VA_ARGS__; } break; \
    }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace mc

using namespace mc;

extern "C" int mc_ln_fwd(const float* x, int64_t x_row_stride, const int32_t* row_index, const float* cls,
                         int64_t cls_period, const float* gamma, const float* beta, void* y, int32_t y_dtype,
                         int64_t y_row_stride, float* mean, float* rstd, int64_t rows, int64_t D, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0) return MC_OK;
    const int vpl = pick_vpl(D);
    MC_CHECK(vpl > 0 && D % 4 == 0 && D > 0, "ln_fwd: D=%lld must be a positive multiple of 4 and <= 1024", (long long)D);
    MC_CHECK(x_row_stride % 4 == 0 && y_row_stride % 4 == 0, "ln_fwd: row strides must be multiples of 4");
    MC_CHECK(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta) && (cls == nullptr || aligned16(cls)),
             "ln_fwd: pointers must be 16-byte aligned");
    MC_CHECK(cls == nullptr || cls_period > 0, "ln_fwd: cls_period must be positive");
    const unsigned grid = (unsigned)ceil_div(rows, kWarpsPerBlock);
    MC_DISPATCH_VPL(vpl, (MC_LAUNCH((ln_fwd_kernel<VPL>), grid, kWarpsPerBlock * 32, 0, stream, 
                             x, x_row_stride, row_index, cls, cls_period, gamma, beta, y, y_dtype, y_row_stride, mean,
                             rstd, rows, (int)D)));
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

extern "C" int mc_ln_bwd(const float* dy, const float* x, int64_t x_row_stride, const int32_t* row_index,
                         const float* cls, int64_t cls_period, const float* mean, const float* rstd, const float* gamma,
                         const float* dres, float* dx, int64_t dx_row_stride, void* dx_act, int32_t act_dtype,
                         float* dgamma, float* dbeta, float* colsum_out, float* rowsum_out, int64_t rowsum_period,
                         float* dcls, int64_t rows, int64_t D, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0) return MC_OK;
    const int vpl = pick_vpl(D);
    MC_CHECK(vpl > 0 && D % 4 == 0 && D > 0, "ln_bwd: D=%lld must be a positive multiple of 4 and <= 1024", (long long)D);
    MC_CHECK(x_row_stride % 4 == 0 && dx_row_stride % 4 == 0, "ln_bwd: row strides must be multiples of 4");
    MC_CHECK(aligned16(dy) && aligned16(x) && aligned16(gamma) && aligned16(dres) && aligned16(dx) && aligned16(dx_act),
             "ln_bwd: pointers must be 16-byte aligned");
    MC_CHECK(rowsum_out == nullptr || rowsum_period > 0, "ln_bwd: rowsum_period must be positive");
    MC_CHECK(cls == nullptr || cls_period > 0, "ln_bwd: cls_period must be positive");
    MC_CHECK(!(dcls != nullptr && colsum_out != nullptr), "ln_bwd: dcls and colsum_out are mutually exclusive");
    // dense block-level case: row kernel + column kernel (see ln_bwd_rows_kernel)
    const bool dense = row_index == nullptr && cls == nullptr && dcls == nullptr && dx != nullptr && x_row_stride == D &&
                       dx_row_stride == D && D % 4 == 0 && aligned16(dgamma) && aligned16(dbeta) && aligned16(colsum_out);
    if (dense) {
        const size_t smem = (size_t)kWarpsPerBlock * 3 * vpl * 128 * sizeof(float);   // <= 96 KB (VPL 8)
        int per_sm = (int)((200 * 1024) / (smem + 1024));
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        int64_t blocks = ceil_div(rows, kWarpsPerBlock);
        const int64_t cap = (int64_t)sm_count() * per_sm;
        if (blocks > cap) blocks = cap;
        MC_DISPATCH_VPL(vpl, {
            static bool attr_set = false;
            if (!attr_set) {
                MC_CUDA(cudaFuncSetAttribute(ln_bwd_rows_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                attr_set = true;
            }
            MC_LAUNCH((ln_bwd_rows_kernel<VPL>), (unsigned)blocks, kWarpsPerBlock * 32, smem, stream, 
                dy, x, mean, rstd, gamma, dres, dx, dx_act, act_dtype, rowsum_out, rowsum_period, dgamma, dbeta, colsum_out,
                rows, (int)D);
        });
        MC_CUDA(cudaGetLastError());
        return MC_OK;
    }
    int64_t blocks = ceil_div(rows, kWarpsPerBlock);
    const int64_t cap = (int64_t)sm_count() * 2;
    if (blocks > cap) blocks = cap;
    MC_DISPATCH_VPL(vpl, (MC_LAUNCH((ln_bwd_kernel<VPL>), (unsigned)blocks, kWarpsPerBlock * 32, 0, stream, 
                             dy, x, x_row_stride, row_index, cls, cls_period, mean, rstd, gamma, dres, dx, dx_row_stride,
                             dx_act, act_dtype, dgamma, dbeta, colsum_out, rowsum_out, rowsum_period, dcls, rows, (int)D)));
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

extern "C" int mc_colsum(const void* x, int32_t dtype, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0 || cols == 0) return MC_OK;
    const int esz = dtype == MC_BF16 ? 2 : 4;
    const bool vec = cols % 8 == 0 && (ld * esz) % 16 == 0 && aligned16(x);
    const int64_t col_blocks = ceil_div(cols, 256);
    int64_t row_blocks = ceil_div((int64_t)sm_count() * 4, col_blocks);
    if (row_blocks > ceil_div(rows, 64)) row_blocks = ceil_div(rows, 64);
    if (row_blocks < 1) row_blocks = 1;
    const int64_t rpb = ceil_div(ceil_div(rows, row_blocks), 8) * 8;
    dim3 grid((unsigned)col_blocks, (unsigned)ceil_div(rows, rpb));
    if (vec) {
        if (dtype == MC_BF16)
            MC_LAUNCH((colsum_vec_kernel<__nv_bfloat16>), grid, 256, 0, stream, reinterpret_cast<const __nv_bfloat16*>(x), rows, cols, ld, out, rpb);
        else
            MC_LAUNCH((colsum_vec_kernel<float>), grid, 256, 0, stream, reinterpret_cast<const float*>(x), rows, cols, ld, out, rpb);
    } else {
        if (dtype == MC_BF16)
            MC_LAUNCH((colsum_kernel<__nv_bfloat16>), grid, 128, 0, stream, reinterpret_cast<const __nv_bfloat16*>(x), rows, cols, ld, out, rpb);
        else
            MC_LAUNCH((colsum_kernel<float>), grid, 128, 0, stream, reinterpret_cast<const float*>(x), rows, cols, ld, out, rpb);
    }
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

extern "C" int mc_rowsum(const void* x, int32_t dtype, int64_t rows, int64_t cols, int64_t ld, int64_t period, float* out,
                         void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0 || cols == 0) return MC_OK;
    MC_CHECK(period > 0, "rowsum: period must be positive");
    const unsigned grid = (unsigned)ceil_div(rows, kWarpsPerBlock);
    const int esz = dtype == MC_BF16 ? 2 : 4;
    const int vec = (cols % 8 == 0 && (ld * esz) % 16 == 0 && aligned16(x)) ? 1 : 0;
    if (dtype == MC_BF16)
        MC_LAUNCH((rowsum_kernel<__nv_bfloat16>), grid, kWarpsPerBlock * 32, 0, stream, reinterpret_cast<const __nv_bfloat16*>(x), rows, cols, ld, period, out, vec);
    else
        MC_LAUNCH((rowsum_kernel<float>), grid, kWarpsPerBlock * 32, 0, stream, reinterpret_cast<const float*>(x), rows, cols, ld, period, out, vec);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

extern "C" int mc_cast_pad(const float* src, int64_t rows, int64_t cols, int64_t src_ld, void* dst, int32_t dst_dtype,
                           int64_t dst_ld, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0 || cols == 0) return MC_OK;
    MC_CHECK(dst_ld >= cols && src_ld >= cols, "cast_pad: leading dimensions smaller than cols");
    const int64_t total = rows * dst_ld;
    MC_LAUNCH((cast_pad_kernel), (unsigned)ceil_div(total, 256), 256, 0, stream, src, rows, cols, src_ld, dst, dst_dtype, dst_ld);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

// Batched transpose of small bf16 matrices: dst[b][c][r] = src[b][r][c] (token-mixing lin1 weights [4P x P] -> W1^T
// [P x 4P], one matrix per Mixer block; the fused token-mixing kernels then fetch BOTH resident weight tiles with plain
// TMA tensor loads instead of a 2-byte gather at every launch).  Pad columns of dst (c_ld > rows) are written as zero.
namespace mc { namespace {
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, int rows, int cols, long long ld_src, long long src_bs,
                      __nv_bfloat16* __restrict__ dst, long long ld_dst, long long dst_bs) {
    MC_PDL_PROLOGUE();
    const __nv_bfloat16* s = src + (long long)blockIdx.y * src_bs;
    __nv_bfloat16* d = dst + (long long)blockIdx.y * dst_bs;
    const long long total = (long long)cols * ld_dst;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / ld_dst), r = (int)(i - (long long)c * ld_dst);       // consecutive threads: consecutive dst elements
        d[i] = r < rows ? s[(long long)r * ld_src + c] : __float2bfloat16_rn(0.f);
    }
}
} }

extern "C" int mc_transpose_bf16(const void* src, int64_t rows, int64_t cols, int64_t ld_src, int64_t src_batch_stride,
                                 void* dst, int64_t ld_dst, int64_t dst_batch_stride, int64_t batch, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0 || cols == 0 || batch == 0) return MC_OK;
    MC_CHECK(ld_src >= cols && ld_dst >= rows, "transpose_bf16: leading dimensions too small");
    MC_CHECK(batch < 65536, "transpose_bf16: too many matrices");
    const int64_t total = cols * ld_dst;
    dim3 grid((unsigned)(ceil_div(total, 256) < 64 ? ceil_div(total, 256) : 64), (unsigned)batch);
    MC_LAUNCH((transpose_bf16_kernel), grid, 256, 0, stream, reinterpret_cast<const __nv_bfloat16*>(src), (int)rows, (int)cols, ld_src,
                                                    src_batch_stride, reinterpret_cast<__nv_bfloat16*>(dst), ld_dst,
                                                    dst_batch_stride);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

extern "C" int mc_im2col(const void* image, int32_t image_is_u8, int64_t B, int64_t R, int64_t patch, void* out,
                         int32_t out_dtype, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0) return MC_OK;
    MC_CHECK(patch > 0 && R % patch == 0, "im2col: resolution %lld not divisible by patch %lld", (long long)R, (long long)patch);
    MC_CHECK(patch % 4 == 0 && (reinterpret_cast<uintptr_t>(image) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "im2col: patch size must be a multiple of 4 and buffers 16-byte aligned");
    const int64_t total = B * 3 * R * R / 4;
    const int64_t blocks = ceil_div(total, 256);
    MC_CHECK(blocks < (1ll << 31), "im2col: too large");
    if (image_is_u8)
        MC_LAUNCH((im2col_kernel<true>), (unsigned)blocks, 256, 0, stream, image, B, (int)R, (int)patch, out, out_dtype);
    else
        MC_LAUNCH((im2col_kernel<false>), (unsigned)blocks, 256, 0, stream, image, B, (int)R, (int)patch, out, out_dtype);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

extern "C" int mc_embed_fwd(const int64_t* text, const float* table, float* x, int64_t B, int64_t C, int64_t W,
                            int64_t vocab, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0) return MC_OK;
    MC_CHECK(W % 4 == 0, "embed: width must be a multiple of 4");
    const int64_t rows = B * C;
    MC_LAUNCH((embed_fwd_kernel), (unsigned)ceil_div(rows, kWarpsPerBlock), kWarpsPerBlock * 32, 0, stream, 
        reinterpret_cast<const long long*>(text), table, x, rows, (int)W, vocab);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

extern "C" int mc_embed_bwd(const int64_t* text, const float* dx, float* dtable, int64_t B, int64_t C, int64_t W,
                            int64_t vocab, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0) return MC_OK;
    const int vpl = pick_vpl(W);
    MC_CHECK(vpl > 0 && W % 4 == 0, "embed_bwd: width must be a multiple of 4 and <= 1024");
    MC_CHECK(aligned16(dtable) && aligned16(dx), "embed_bwd: pointers must be 16-byte aligned");
    const unsigned grid = (unsigned)ceil_div(B, kWarpsPerBlock);
    MC_DISPATCH_VPL(vpl, (MC_LAUNCH((embed_bwd_kernel<VPL>), grid, kWarpsPerBlock * 32, 0, stream, 
                             reinterpret_cast<const long long*>(text), dx, dtable, B, (int)C, (int)W, vocab)));
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

extern "C" int mc_eot_rows(const int64_t* text, int32_t* eot_row, int64_t B, int64_t C, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0) return MC_OK;
    MC_CHECK(B * C < (1ll << 31), "eot_rows: batch too large");
    MC_LAUNCH((eot_rows_kernel), (unsigned)ceil_div(B, 128), 128, 0, stream, reinterpret_cast<const long long*>(text), eot_row, B, (int)C);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

extern "C" int mc_l2norm_fwd(const float* f, float* u, float* inv_norm, int64_t rows, int64_t E, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0) return MC_OK;
    const int vpl = pick_vpl(E);
    MC_CHECK(vpl > 0 && E % 4 == 0, "l2norm: E must be a multiple of 4 and <= 1024");
    const unsigned grid = (unsigned)ceil_div(rows, kWarpsPerBlock);
    MC_DISPATCH_VPL(vpl, (MC_LAUNCH((l2norm_fwd_kernel<VPL>), grid, kWarpsPerBlock * 32, 0, stream, f, u, inv_norm, rows, (int)E)));
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

extern "C" int mc_l2norm_bwd(const float* du, const float* u, const float* inv_norm, float* df, void* df_act,
                             int32_t act_dtype, int64_t rows, int64_t E, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0) return MC_OK;
    const int vpl = pick_vpl(E);
    MC_CHECK(vpl > 0 && E % 4 == 0, "l2norm: E must be a multiple of 4 and <= 1024");
    const unsigned grid = (unsigned)ceil_div(rows, kWarpsPerBlock);
    MC_DISPATCH_VPL(vpl, (MC_LAUNCH((l2norm_bwd_kernel<VPL>), grid, kWarpsPerBlock * 32, 0, stream, du, u, inv_norm, df, df_act,
                                                                                          act_dtype, rows, (int)E)));
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}
