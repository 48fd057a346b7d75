// Fused optimizer step over flat fp32 buffers: global gradient norm (clip_grad_norm_(.., 20),
// training/training.py:181) and AdamW with the reference's two parameter groups
// (training/training.py:66-82,185): lr 5e-4, betas (0.9, 0.98), eps 1e-6, weight decay 0.2 on
// tensors with ndim >= 2 whose name has no "bn"/"ln"/"bias"/"logit_scale", 0 elsewhere.
// HBM-bound: 16 B read + 12 B (+2 B bf16 mirror) written per parameter, one launch for the model.
#include "common.cuh"

namespace mc {
namespace {

// Rank-consistent (deterministic) global norm: every block writes its partial sum into a fixed slot and the LAST block
// to finish (ticket counter) adds the slots in index order, so all data-parallel replicas - which hold bit-identical
// all-reduced gradients - compute bit-identical clip coefficients (torch's clip_grad_norm_ is deterministic too).
__device__ float g_sumsq_partials[4096];
__device__ unsigned int g_sumsq_ticket = 0;

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
    MC_PDL_PROLOGUE();
    __shared__ float red[8];
    __shared__ bool last;
    float s = 0.f;
    const long long n4 = n / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0) for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        s = red[threadIdx.x];
        s += __shfl_xor_sync(0xffu, s, 4);
        s += __shfl_xor_sync(0xffu, s, 2);
        s += __shfl_xor_sync(0xffu, s, 1);
        if (threadIdx.x == 0) {
            g_sumsq_partials[blockIdx.x] = s;
            __threadfence();
            last = atomicAdd(&g_sumsq_ticket, 1u) == gridDim.x - 1;
        }
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // fixed-order tree over the slots (same order on every launch and every rank)
    float t = 0.f;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) t += __ldcg(&g_sumsq_partials[i]);
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < 8; ++w) tot += red[w];
        out[0] += tot;
        g_sumsq_ticket = 0;
    }
}

// Per-step scalars on the DEVICE (graph-capturable, no pinned host slot to race on): state = {Adam step count t,
// scheduler step s}; writes hyper = {lr(s), 1 - beta1^(t+1), 1 - beta2^(t+1)} and advances both counters.  The
// schedule is CosineAnnealingWarmupRestarts with cycle_mult = 1, gamma = 1 (training/training.py:83-89), restated
// in optim.py::cosine_warmup_lr; fixed_lr >= 0 bypasses it.
__global__ void sched_step_kernel(long long* __restrict__ state, float* __restrict__ hyper, long long first_cycle_steps,
                                  double max_lr, double min_lr, long long warmup_steps, double beta1, double beta2,
                                  double fixed_lr) {
    MC_PDL_PROLOGUE();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long t = state[0] + 1, s = state[1];
    double lr = fixed_lr;
    if (fixed_lr < 0.0) {
        const long long in_cycle = s % first_cycle_steps;
        if (in_cycle < warmup_steps) lr = (max_lr - min_lr) * double(in_cycle) / double(warmup_steps) + min_lr;
        else
            lr = min_lr + (max_lr - min_lr) *
                              (1.0 + cos(3.14159265358979323846 * double(in_cycle - warmup_steps) /
                                         double(first_cycle_steps - warmup_steps))) / 2.0;
    }
    hyper[0] = (float)lr;
    hyper[1] = (float)(1.0 - pow(beta1, (double)t));
    hyper[2] = (float)(1.0 - pow(beta2, (double)t));
    state[0] = t;
    state[1] = s + 1;
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ pb, const uint8_t* __restrict__ decay_flags, long long n4,
             const float* __restrict__ sumsq, const float* __restrict__ hyper, float grad_mul, float max_norm,
             float beta1, float beta2, float eps, float weight_decay) {
    MC_PDL_PROLOGUE();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float lr = hyper[0], bc1 = hyper[1], bc2 = hyper[2];
    float gs = grad_mul;
    if (sumsq != nullptr && max_norm > 0.f) {
        const float norm = sqrtf(sumsq[0]) * grad_mul;
        const float coef = max_norm / (norm + 1e-6f);
        if (coef < 1.0f) gs *= coef;
    }
    const float decay = (decay_flags != nullptr && decay_flags[i / 16]) ? 1.0f - lr * weight_decay : 1.0f;
    const float step = lr / bc1, rs2 = rsqrtf(bc2);
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
#define MC_ADAM1(c)                                              \
    {                                                            \
        const float gg = gv.c * gs;                              \
        mv.c = beta1 * mv.c + (1.0f - beta1) * gg;               \
        vv.c = beta2 * vv.c + (1.0f - beta2) * gg * gg;          \
        pv.c = pv.c * decay - step * mv.c / (sqrtf(vv.c) * rs2 + eps); \
    }
    MC_ADAM1(x) MC_ADAM1(y) MC_ADAM1(z) MC_ADAM1(w)
#undef MC_ADAM1
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (pb != nullptr) {
        uint2 pk;
        pk.x = pack_bf16x2(pv.x, pv.y);
        pk.y = pack_bf16x2(pv.z, pv.w);
        reinterpret_cast<uint2*>(pb)[i] = pk;
    }
}

}  // namespace
}  // namespace mc

using namespace mc;

extern "C" int mc_sumsq(const float* g, int64_t n, float* out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return MC_OK;
    MC_CHECK((reinterpret_cast<uintptr_t>(g) & 15) == 0, "sumsq: pointer must be 16-byte aligned");
    int64_t blocks = ceil_div(n / 4 + 1, 256);
    int64_t cap = (int64_t)sm_count() * 8;
    if (cap > 4096) cap = 4096;
    if (blocks > cap) blocks = cap;
    MC_LAUNCH((sumsq_kernel), (unsigned)blocks, 256, 0, stream, g, n, out);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

extern "C" int mc_sched_step(int64_t* state, float* hyper, int64_t first_cycle_steps, double max_lr, double min_lr,
                             int64_t warmup_steps, double beta1, double beta2, double fixed_lr, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MC_CHECK(state != nullptr && hyper != nullptr, "sched_step: null state / hyper");
    MC_CHECK(fixed_lr >= 0.0 || first_cycle_steps > warmup_steps, "sched_step: first_cycle_steps must exceed warmup_steps");
    MC_LAUNCH((sched_step_kernel), 1, 32, 0, stream, reinterpret_cast<long long*>(state), hyper, first_cycle_steps, max_lr, min_lr,
                                            warmup_steps, beta1, beta2, fixed_lr);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

extern "C" int mc_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, const uint8_t* decay_flags, int64_t n,
                        const float* sumsq, const float* hyper, float grad_mul, float max_norm, float beta1, float beta2,
                        float eps, float weight_decay, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return MC_OK;
    MC_CHECK(n % 64 == 0, "adamw: flat buffers are padded to 64-element chunks (n=%lld)", (long long)n);
    MC_CHECK(hyper != nullptr, "adamw: hyper (device {lr, 1-b1^t, 1-b2^t}) is required");
    const int64_t n4 = n / 4;
    MC_LAUNCH((adamw_kernel), (unsigned)ceil_div(n4, 256), 256, 0, stream, p, g, m, v, reinterpret_cast<__nv_bfloat16*>(p_bf16),
                                                                  decay_flags, n4, sumsq, hyper, grad_mul, max_norm, beta1,
                                                                  beta2, eps, weight_decay);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}
