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

}  // namespace
}  // namespace mc

using namespace mc;

extern "C" int64_t mc_head_workspace_bytes(int64_t n, int64_t N, int64_t E) {
    const int sp = head_splits(n, N, sm_count());
    return 2ll * sp * n * (E + 3) * (int64_t)sizeof(float);
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
