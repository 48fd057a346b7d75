// Contrastive head: loss + gradients of training/training.py:158-168 (and the autograd backward
// of :170) without ever writing the [n x N] logit matrices.
//
// Because the gathered features are detached (training.py:158-159), each direction is a row
// soft-max cross-entropy of s * U_loc @ V_all^T whose gradient wrt the local rows is
//     dU_loc = s/(2n) * (softmax(A) @ V_all - V_all[g]),      g_i = rank*n + i
// i.e. loss AND gradients come out of one flash-attention-forward-shaped pass per direction
// (Q = local rows, K = V = gathered rows, head dim E, scale s).
//
// Pass 1 (head_partial): grid (row blocks of 32, column splits, 2 directions); online softmax over
//   32-column tiles; unnormalised O, running max m, running sum l and the target logit go to the
//   workspace.  Pass 2 (head_combine): merges the splits, emits loss, dU, d(log scale).
// All fp32 FFMA (the reference computes the head in fp32 outside autocast).
#include "common.cuh"

namespace mc {
namespace {

constexpr int BR = 32, BC = 32, kHeadThreads = 256, kMaxE = 512, kColsPerLane = kMaxE / 32;

__host__ __device__ inline int head_splits(long long n, long long N, int sms) {
    const long long rb = (n + BR - 1) / BR;
    long long sp = (2ll * sms + 2 * rb - 1) / (2 * rb);
    const long long max_sp = (N + BC - 1) / BC;
    if (sp > max_sp) sp = max_sp;
    if (sp < 1) sp = 1;
    return (int)sp;
}

// workspace layout per (dir, split): m[n], l[n], tgt[n], O[n][E]
__global__ void __launch_bounds__(kHeadThreads)
head_partial_kernel(const float* __restrict__ ui, const float* __restrict__ ut, const float* __restrict__ ui_all,
                    const float* __restrict__ ut_all, const float* __restrict__ log_scale, long long n, long long N,
                    int E, long long rank, int splits, float* __restrict__ ws) {
    MC_PDL_PROLOGUE();
    extern __shared__ float sm[];
    const int ldq = E + 4;
    float* Qs = sm;                 // [BR][ldq]
    float* Ks = Qs + BR * ldq;      // [BC][ldq]
    float* Ps = Ks + BC * ldq;      // [BR][BC+1]
    const int dir = blockIdx.z, split = blockIdx.y;
    const float* Q = dir == 0 ? ui : ut;
    const float* KV = dir == 0 ? ut_all : ui_all;
    const long long r0 = (long long)blockIdx.x * BR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float s = expf(log_scale[0]);

    // Q tile
    for (int i = threadIdx.x; i < BR * (E / 4); i += kHeadThreads) {
        const int r = i / (E / 4), c = (i % (E / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < n) v = *reinterpret_cast<const float4*>(Q + (r0 + r) * E + c);
        *reinterpret_cast<float4*>(Qs + r * ldq + c) = v;
    }

    float m_run[4], l_run[4], tgt[4];
    float O[4][kColsPerLane];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_run[i] = -INFINITY;
        l_run[i] = 0.f;
        tgt[i] = 0.f;
#pragma unroll
        for (int k = 0; k < kColsPerLane; ++k) O[i][k] = 0.f;
    }
    const int ncl = (E + 31) / 32;  // columns per lane actually used

    const long long tiles = (N + BC - 1) / BC;
    const long long t_begin = tiles * split / splits, t_end = tiles * (split + 1) / splits;
    for (long long t = t_begin; t < t_end; ++t) {
        const long long j0 = t * BC;
        __syncthreads();  // previous tile's Ks fully consumed (and Qs written, first iteration)
        for (int i = threadIdx.x; i < BC * (E / 4); i += kHeadThreads) {
            const int r = i / (E / 4), c = (i % (E / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j0 + r < N) v = *reinterpret_cast<const float4*>(KV + (j0 + r) * E + c);
            *reinterpret_cast<float4*>(Ks + r * ldq + c) = v;
        }
        __syncthreads();
        // S[r][lane] for the warp's 4 rows
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* kp = Ks + lane * ldq;
        for (int c = 0; c < E; c += 4) {
            const float4 kv = *reinterpret_cast<const float4*>(kp + c);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 qv = *reinterpret_cast<const float4*>(Qs + (warp * 4 + i) * ldq + c);
                acc[i] = fmaf(qv.x, kv.x, acc[i]);
                acc[i] = fmaf(qv.y, kv.y, acc[i]);
                acc[i] = fmaf(qv.z, kv.z, acc[i]);
                acc[i] = fmaf(qv.w, kv.w, acc[i]);
            }
        }
        const bool col_ok = j0 + lane < N;
        float alpha[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long row = r0 + warp * 4 + i;
            const float sv = col_ok ? s * acc[i] : -INFINITY;
            const long long label = rank * n + row;
            tgt[i] += warp_sum((col_ok && j0 + lane == label) ? sv : 0.f);
            const float m_new = fmaxf(m_run[i], warp_max(sv));
            const float p = col_ok ? expf(sv - m_new) : 0.f;
            alpha[i] = (m_run[i] == -INFINITY) ? 0.f : expf(m_run[i] - m_new);
            l_run[i] = l_run[i] * alpha[i] + warp_sum(p);
            m_run[i] = m_new;
            Ps[(warp * 4 + i) * (BC + 1) + lane] = p;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < kColsPerLane; ++k) O[i][k] *= alpha[i];
        for (int jj = 0; jj < BC; ++jj) {
            float pv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pv[i] = Ps[(warp * 4 + i) * (BC + 1) + jj];
#pragma unroll
            for (int k = 0; k < kColsPerLane; ++k) {
                if (k < ncl) {
                    const int c = lane + 32 * k;
                    const float kv = c < E ? Ks[jj * ldq + c] : 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) O[i][k] = fmaf(pv[i], kv, O[i][k]);
                }
            }
        }
        __syncwarp();
    }

    // partials
    float* base = ws + ((long long)(dir * splits + split)) * n * (E + 3);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long row = r0 + warp * 4 + i;
        if (row >= n) continue;
        if (lane == 0) {
            base[row] = m_run[i];
            base[n + row] = l_run[i];
            base[2 * n + row] = tgt[i];
        }
        float* op = base + 3 * n + row * E;
#pragma unroll
        for (int k = 0; k < kColsPerLane; ++k) {
            const int c = lane + 32 * k;
            if (k < ncl && c < E) op[c] = O[i][k];
        }
    }
}

// one warp per (dir, row)
__global__ void __launch_bounds__(256)
head_combine_kernel(const float* __restrict__ ui, const float* __restrict__ ut, const float* __restrict__ ui_all,
                    const float* __restrict__ ut_all, const float* __restrict__ log_scale, long long n, long long N,
                    int E, long long rank, int splits, const float* __restrict__ ws, float grad_scale,
                    float* __restrict__ loss, float* __restrict__ dui, float* __restrict__ dut,
                    float* __restrict__ dlog_scale) {
    MC_PDL_PROLOGUE();
    const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= 2 * n) return;
    const int dir = (int)(wid / n);
    const long long row = wid % n;
    const float* Q = dir == 0 ? ui : ut;
    const float* KV = dir == 0 ? ut_all : ui_all;
    float* dQ = dir == 0 ? dui : dut;
    const float s = expf(log_scale[0]);
    const long long stride = n * (long long)(E + 3);
    const float* base = ws + (long long)dir * splits * stride;

    float M = -INFINITY;
    for (int sp = 0; sp < splits; ++sp) M = fmaxf(M, base[sp * stride + row]);
    float L = 0.f, tg = 0.f;
    for (int sp = 0; sp < splits; ++sp) {
        const float m = base[sp * stride + row];
        const float w = (m == -INFINITY) ? 0.f : expf(m - M);
        L += base[sp * stride + n + row] * w;
        tg += base[sp * stride + 2 * n + row];
    }
    const float lse = M + logf(L);
    const float inv2n = 0.5f / (float)n;
    const long long label = rank * n + row;
    float qo = 0.f;
    for (int c = lane; c < E; c += 32) {
        float o = 0.f;
        for (int sp = 0; sp < splits; ++sp) {
            const float m = base[sp * stride + row];
            const float w = (m == -INFINITY) ? 0.f : expf(m - M);
            o += base[sp * stride + 3 * n + row * E + c] * w;
        }
        o /= L;
        qo += Q[row * E + c] * o;
        dQ[row * E + c] = grad_scale * s * inv2n * (o - KV[label * E + c]);
    }
    qo = warp_sum(qo);
    if (lane == 0) {
        atomicAdd(loss, (lse - tg) * inv2n);
        atomicAdd(dlog_scale, grad_scale * (s * qo - tg) * inv2n);
    }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core path for global-batch sizes (BASELINE configs[2]: n = 4096 local rows against N = 32768 gathered rows,
// 0.55 TFLOP per rank and step - tens of milliseconds on FFMA).  Both contractions of a direction run on the tcgen05
// GEMM engine with fp32-class accuracy through a bf16 x 3 split (x = hi + lo, x.y ~ hi.hi + hi.lo + lo.hi, relative
// error 2^-16 per product): the three terms are ONE GEMM whose reduction also runs over a "batch" of 3 operand pieces
// (k_spans_batch), pieces (Qh, Qh, Ql) against (Kh, Kl, Kh).  The logits exist only as one [n x 4096] fp32 slab at a
// time (64 MB, L2-sized: written by the S GEMM, read once by the softmax kernel, overwritten by the next slab), never
// as the [n x N] matrix the reference materialises twice (training.py:162-163):
//     per direction:  split Q, K                       head_split_kernel
//       per slab j:   S      = Q' K'_j^T               mc_gemm_bf16_tc (K = E x 3)
//                     online softmax over the slab: m, l, target logit, P~ = exp(s S - m) as (Ph, Ph, Pl), O *= alpha
//                     O     += P' K'_j                  mc_gemm_bf16_tc (K = 4096 x 3, accumulate, split-K)
//                     loss, dQ, d(log scale)            head_finalize_kernel   (same formulas as head_combine_kernel)
// ------------------------------------------------------------------------------------------------
constexpr int kSlab = 4096;

// dst[piece][row][e]: mode 0 (queries) = (hi, hi, lo), mode 1 (keys) = (hi, lo, hi)
__global__ void __launch_bounds__(256)
head_split_kernel(const float* __restrict__ src, long long rows, int E, int mode, __nv_bfloat16* __restrict__ dst) {
    MC_PDL_PROLOGUE();
    const long long total = rows * E / 4;
    const long long piece = rows * (long long)E;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        const float x[4] = {v.x, v.y, v.z, v.w};
        uint32_t hi[2], lo[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * k]), h1 = __float2bfloat16_rn(x[2 * k + 1]);
            hi[k] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
            lo[k] = pack_bf16x2(x[2 * k] - __bfloat162float(h0), x[2 * k + 1] - __bfloat162float(h1));
        }
        uint2* d = reinterpret_cast<uint2*>(dst) + i;
        const uint2 H = make_uint2(hi[0], hi[1]), L = make_uint2(lo[0], lo[1]);
        d[0] = H;
        d[piece / 4] = mode == 0 ? H : L;
        d[2 * (piece / 4)] = mode == 0 ? L : H;
    }
}

__global__ void head_fill_kernel(float* __restrict__ p, long long n, float value) {
    MC_PDL_PROLOGUE();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = value;
}

// one block per local row: online-softmax update with the logits of one slab
__global__ void __launch_bounds__(256)
head_slab_kernel(const float* __restrict__ S, int ldS, long long n, int cols, long long j0, long long rank,
                 const float* __restrict__ log_scale, float* __restrict__ m_run, float* __restrict__ l_run,
                 float* __restrict__ tgt, float* __restrict__ O, int E, __nv_bfloat16* __restrict__ P) {
    MC_PDL_PROLOGUE();
    __shared__ float red[8];
    __shared__ float bc[2];
    const long long row = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float s = expf(log_scale[0]);
    const float* sr = S + row * (long long)ldS;
    constexpr int kPer = kSlab / 256;   // 16 logits per thread, strided by 256: coalesced
    float v[kPer];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int c = tid + 256 * k;
        v[k] = c < cols ? s * sr[c] : -INFINITY;
        mx = fmaxf(mx, v[k]);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        float a = red[0];
        for (int w = 1; w < 8; ++w) a = fmaxf(a, red[w]);
        const float m_old = m_run[row];
        const float m_new = fmaxf(m_old, a);
        bc[0] = m_new;
        bc[1] = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
        m_run[row] = m_new;
    }
    __syncthreads();
    const float m_new = bc[0], alpha = bc[1];
    const long long label = rank * n + row;
    float sum = 0.f;
    const long long piece = n * (long long)kSlab;
    __nv_bfloat16* pr = P + row * (long long)kSlab;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int c = tid + 256 * k;
        float p = 0.f;
        if (c < cols) {
            p = expf(v[k] - m_new);
            if (j0 + c == label) tgt[row] = v[k];
        }
        sum += p;
        const __nv_bfloat16 h = __float2bfloat16_rn(p);
        const __nv_bfloat16 l = __float2bfloat16_rn(p - __bfloat162float(h));
        pr[c] = h;
        pr[piece + c] = h;
        pr[2 * piece + c] = l;
    }
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += red[w];
        l_run[row] = l_run[row] * alpha + a;
    }
    // rescale the running output row before this slab's P~ K is accumulated into it
    float* orow = O + row * (long long)E;
    for (int c = tid; c < E; c += 256) orow[c] *= alpha;
}

// one warp per local row (same formulas as head_combine_kernel, one "split")
__global__ void __launch_bounds__(256)
head_finalize_kernel(const float* __restrict__ Q, const float* __restrict__ KV, const float* __restrict__ O,
                     const float* __restrict__ m_run, const float* __restrict__ l_run, const float* __restrict__ tgt,
                     const float* __restrict__ log_scale, long long n, int E, long long rank, float grad_scale,
                     float* __restrict__ loss, float* __restrict__ dQ, float* __restrict__ dlog_scale) {
    MC_PDL_PROLOGUE();
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float s = expf(log_scale[0]);
    const float L = l_run[row], lse = m_run[row] + logf(L), tg = tgt[row];
    const float inv2n = 0.5f / (float)n;
    const long long label = rank * n + row;
    float qo = 0.f;
    for (int c = lane; c < E; c += 32) {
        const float o = O[row * E + c] / L;
        qo += Q[row * E + c] * o;
        dQ[row * E + c] = grad_scale * s * inv2n * (o - KV[label * E + c]);
    }
    qo = warp_sum(qo);
    if (lane == 0) {
        atomicAdd(loss, (lse - tg) * inv2n);
        atomicAdd(dlog_scale, grad_scale * (s * qo - tg) * inv2n);
    }
}

inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

struct HeadTcPlan {
    int64_t off_q, off_k, off_s, off_p, off_o, off_stats, total;
};
inline HeadTcPlan head_tc_plan(int64_t n, int64_t N, int64_t E) {
    HeadTcPlan p{};
    int64_t off = 0;
    p.off_q = off; off += align256(3 * n * E * 2);
    p.off_k = off; off += align256(3 * N * E * 2);
    p.off_s = off; off += align256(n * (int64_t)kSlab * 4);
    p.off_p = off; off += align256(3 * n * (int64_t)kSlab * 2);
    p.off_o = off; off += align256(n * E * 4);
    p.off_stats = off; off += align256(3 * n * 4);
    p.total = off;
    return p;
}

// MC_HEAD_TC: "auto" (default: tensor-core path from 2^24 logits per direction), "0" (always FFMA), "1" (whenever the shape allows)
inline bool head_use_tc(int64_t n, int64_t N, int64_t E) {
    const char* v = getenv("MC_HEAD_TC");       // read per call (tests switch it)
    const int mode = (v == nullptr || v[0] == 'a' || v[0] == '\0') ? 2 : (atoi(v) != 0 ? 1 : 0);
    const bool shape_ok = E % 64 == 0 && E <= kMaxE && n >= 128 && N >= 128;
    if (mode == 0 || !shape_ok) return false;
    if (mode == 1) return true;
    return n * N >= (1ll << 24);
}

int head_tc_direction(const float* Q, const float* KV, const float* log_scale, int64_t n, int64_t N, int64_t E, int64_t rank,
                      float grad_scale, float* loss, float* dQ, float* dlog_scale, uint8_t* ws, cudaStream_t stream) {
    const HeadTcPlan pl = head_tc_plan(n, N, E);
    __nv_bfloat16* q3 = reinterpret_cast<__nv_bfloat16*>(ws + pl.off_q);
    __nv_bfloat16* k3 = reinterpret_cast<__nv_bfloat16*>(ws + pl.off_k);
    float* S = reinterpret_cast<float*>(ws + pl.off_s);
    __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(ws + pl.off_p);
    float* O = reinterpret_cast<float*>(ws + pl.off_o);
    float* m_run = reinterpret_cast<float*>(ws + pl.off_stats);
    float* l_run = m_run + n;
    float* tgt = l_run + n;
    const int sms = sm_count();
    MC_LAUNCH((head_split_kernel), (unsigned)(sms * 4), 256, 0, stream, Q, (long long)n, (int)E, 0, q3);
    MC_LAUNCH((head_split_kernel), (unsigned)(sms * 4), 256, 0, stream, KV, (long long)N, (int)E, 1, k3);
    MC_CUDA(cudaGetLastError());
    MC_CUDA(cudaMemsetAsync(O, 0, (size_t)n * E * 4, stream));
    MC_CUDA(cudaMemsetAsync(l_run, 0, (size_t)n * 4, stream));
    MC_CUDA(cudaMemsetAsync(tgt, 0, (size_t)n * 4, stream));
    MC_LAUNCH((head_fill_kernel), (unsigned)ceil_div(n, 256), 256, 0, stream, m_run, (long long)n, -INFINITY);
    MC_CUDA(cudaGetLastError());
    for (int64_t j0 = 0; j0 < N; j0 += kSlab) {
        const int64_t cols = N - j0 < kSlab ? N - j0 : kSlab;
        mc_gemm_params g{};
        // S[n x cols] = sum over the 3 pieces of  Q'_piece [n x E] . K'_piece[j0 .. j0 + cols)^T
        g.M = n; g.N = cols; g.K = E; g.batch = 3;
        g.A = q3; g.a_major = MC_MAJOR_K; g.lda = E; g.a_batch_stride = n * E;
        g.B = k3 + j0 * E; g.b_major = MC_MAJOR_K; g.ldb = E; g.b_batch_stride = N * E;
        g.k_spans_batch = 1;
        g.C = S; g.c_dtype = MC_F32; g.ldc = kSlab; g.c_batch_stride = 0;
        g.split_k = 1;
        int rc = mc_gemm_bf16_tc(&g, stream);
        if (rc != MC_OK) return rc;
        MC_LAUNCH((head_slab_kernel), (unsigned)n, 256, 0, stream, (const float*)S, (int)kSlab, (long long)n, (int)cols,
                  (long long)j0, (long long)rank, log_scale, m_run, l_run, tgt, O, (int)E, P);
        MC_CUDA(cudaGetLastError());
        // O[n x E] += sum over the 3 pieces of  P'_piece [n x cols] . K'_piece[j0 .. j0 + cols)   (keys as the MN-major operand)
        mc_gemm_params h{};
        h.M = n; h.N = E; h.K = cols; h.batch = 3;
        h.A = P; h.a_major = MC_MAJOR_K; h.lda = kSlab; h.a_batch_stride = n * (int64_t)kSlab;
        h.B = k3 + j0 * E; h.b_major = MC_MAJOR_MN; h.ldb = E; h.b_batch_stride = N * E;
        h.k_spans_batch = 1;
        h.C = O; h.c_dtype = MC_F32; h.ldc = E; h.c_batch_stride = 0;
        h.accumulate = 1; h.split_k = 0;
        rc = mc_gemm_bf16_tc(&h, stream);
        if (rc != MC_OK) return rc;
    }
    MC_LAUNCH((head_finalize_kernel), (unsigned)ceil_div(n, 8), 256, 0, stream, Q, KV, (const float*)O, (const float*)m_run,
              (const float*)l_run, (const float*)tgt, log_scale, (long long)n, (int)E, (long long)rank, grad_scale, loss, dQ,
              dlog_scale);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

}  // namespace
}  // namespace mc

using namespace mc;

extern "C" int64_t mc_head_workspace_bytes(int64_t n, int64_t N, int64_t E) {
    const int sp = head_splits(n, N, sm_count());
    const int64_t simt = 2ll * sp * n * (E + 3) * (int64_t)sizeof(float);
    if (!head_use_tc(n, N, E)) return simt;
    const int64_t tc = head_tc_plan(n, N, E).total;
    return tc > simt ? tc : simt;
}

extern "C" int mc_head_fwd_bwd(const float* ui, const float* ut, const float* ui_all, const float* ut_all,
                               const float* log_scale, int64_t n, int64_t N, int64_t E, int64_t rank, float grad_scale,
                               float* loss, float* dui, float* dut, float* dlog_scale, void* workspace,
                               int64_t workspace_bytes, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MC_CHECK(n > 0 && N > 0 && E > 0, "head: empty problem");
    MC_CHECK(E % 4 == 0 && E <= kMaxE, "head: E=%lld must be a multiple of 4 and <= %d", (long long)E, kMaxE);
    MC_CHECK((rank + 1) * n <= N, "head: labels rank*n+i exceed N (rank=%lld n=%lld N=%lld)", (long long)rank,
             (long long)n, (long long)N);
    const int sp = head_splits(n, N, sm_count());
    MC_CHECK(workspace != nullptr && workspace_bytes >= mc_head_workspace_bytes(n, N, E), "head: workspace too small");
    MC_CHECK((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "head: workspace must be 256-byte aligned");
    if (head_use_tc(n, N, E)) {
        uint8_t* wsb = reinterpret_cast<uint8_t*>(workspace);
        int rc = head_tc_direction(ui, ut_all, log_scale, n, N, E, rank, grad_scale, loss, dui, dlog_scale, wsb, stream);
        if (rc != MC_OK) return rc;
        return head_tc_direction(ut, ui_all, log_scale, n, N, E, rank, grad_scale, loss, dut, dlog_scale, wsb, stream);
    }
    const size_t smem = (size_t)((BR + BC) * (E + 4) + BR * (BC + 1)) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        MC_CUDA(cudaFuncSetAttribute(head_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)ceil_div(n, BR), (unsigned)sp, 2);
    MC_LAUNCH((head_partial_kernel), grid, kHeadThreads, smem, stream, ui, ut, ui_all, ut_all, log_scale, n, N, (int)E, rank, sp,
                                                              reinterpret_cast<float*>(workspace));
    MC_CUDA(cudaGetLastError());
    MC_LAUNCH((head_combine_kernel), (unsigned)ceil_div(2 * n, 8), 256, 0, stream, ui, ut, ui_all, ut_all, log_scale, n, N, (int)E,
                                                                         rank, sp, reinterpret_cast<const float*>(workspace),
                                                                         grad_scale, loss, dui, dut, dlog_scale);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}
