// SIMT FFMA GEMM engine: fp32 operands, fp32 accumulate, same mc_gemm_params contract as the tensor
// core engine.  This is the "fp32 1e-5" validation precision of BASELINE.json's north_star (tcgen05
// has no fp32 MMA; plain TF32 is ~1e-3).  It is a precision mode of the CUDA product path, not a
// fallback: it runs on the GPU only and is selected explicitly (precision="fp32").
// Same call sites as gemm_tc.cu (training/clip/model.py:206-222,258,272,288,424).
#include "common.cuh"

namespace mc {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtArgs {
    int M, N, K, batch, k_spans_batch, out_batch;
    const float* A; long long a_sm, a_sk, a_bs;  // element strides for (m, k, batch)
    const float* B; long long b_sn, b_sk, b_bs;
    int a_kfast, b_kfast;
    float* C; long long ldc, c_bs;
    int accumulate, row_remap;
    const float* bias; int bias_mode;
    float* zout; long long ldz, z_bs;
    const float* zin; long long ldzin, zin_bs;
    int act;
    const float* R; long long ldr, r_bs;
    float* rowsum_out;
    int c_transposed;
    int splits, k_per_split;   // split-K over blockIdx.z (plain epilogue only): partial sums leave through atomicAdd
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtArgs g) {
    MC_PDL_PROLOGUE();
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tiles_m = (g.M + TM - 1) / TM;
    const int tm = blockIdx.x % tiles_m, tn = blockIdx.x / tiles_m;
    const int ob = blockIdx.y;
    const int m0 = tm * TM, n0 = tn * TN;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;  // thread owns rows ty*4.., cols tx*4..

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nb = g.k_spans_batch ? g.batch : 1;
    for (int bi = 0; bi < nb; ++bi) {
        const int bb = g.k_spans_batch ? bi : ob;
        const float* Ab = g.A + (long long)bb * g.a_bs;
        const float* Bb = g.B + (long long)bb * g.b_bs;
        const int k_begin = g.splits > 1 ? (int)blockIdx.z * g.k_per_split : 0;
        const int k_end = g.splits > 1 ? (k_begin + g.k_per_split < g.K ? k_begin + g.k_per_split : g.K) : g.K;
        for (int k0 = k_begin; k0 < k_end; k0 += TK) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                int r, k;
                if (g.a_kfast) { k = tid % TK; r = tid / TK + it * 16; } else { r = tid % TM; k = tid / TM + it * 4; }
                const int gm = m0 + r, gk = k0 + k;
                As[k][r] = (gm < g.M && gk < k_end) ? Ab[(long long)gm * g.a_sm + (long long)gk * g.a_sk] : 0.f;
                if (g.b_kfast) { k = tid % TK; r = tid / TK + it * 16; } else { r = tid % TN; k = tid / TN + it * 4; }
                const int gn = n0 + r;
                const int gk2 = k0 + k;
                Bs[k][r] = (gn < g.N && gk2 < k_end) ? Bb[(long long)gn * g.b_sn + (long long)gk2 * g.b_sk] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
        const long long crow = g.row_remap > 0 ? (long long)m + m / g.row_remap + 1 : (long long)m;
        float rsum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float x = acc[i][j];
            if (g.bias_mode == MC_BIAS_N) x += g.bias[n];
            else if (g.bias_mode == MC_BIAS_M) x += g.bias[m];
            if (g.zout) g.zout[(long long)ob * g.z_bs + crow * g.ldz + n] = x;
            if (g.act == MC_ACT_GELU) x = quick_gelu_precise(x);
            else if (g.act == MC_ACT_GELU_BWD) x *= quick_gelu_grad_precise(g.zin[(long long)ob * g.zin_bs + crow * g.ldzin + n]);
            if (g.R) x += g.c_transposed ? g.R[(long long)ob * g.r_bs + (long long)n * g.ldr + crow]
                                         : g.R[(long long)ob * g.r_bs + crow * g.ldr + n];
            float* cp = g.c_transposed ? g.C + (long long)ob * g.c_bs + (long long)n * g.ldc + crow
                                       : g.C + (long long)ob * g.c_bs + crow * g.ldc + n;
            if (g.splits > 1) atomicAdd(cp, x);
            else *cp = g.accumulate ? *cp + x : x;
            rsum += x;
        }
        if (g.rowsum_out != nullptr) atomicAdd(g.rowsum_out + m, rsum);
    }
}

}  // namespace
}  // namespace mc

using namespace mc;

extern "C" int mc_gemm_f32_simt(const mc_gemm_params* p, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MC_CHECK(p != nullptr, "null params");
    MC_CHECK(p->M > 0 && p->N > 0 && p->K > 0 && p->batch > 0, "gemm: empty problem");
    MC_CHECK(p->A && p->B && p->C, "gemm: null operand");
    MC_CHECK(p->c_dtype == MC_F32, "simt gemm writes fp32 only");
    MC_CHECK(p->A2 == nullptr && p->B2 == nullptr, "simt gemm: the recompute operand pair (A2/B2) is a tensor-core engine feature");
    MC_CHECK(p->rowstat_out == nullptr, "simt gemm: rowstat_out is a tensor-core engine feature");
    MC_CHECK(p->act != MC_ACT_GELU_BWD || p->zin != nullptr, "gemm: GELU_BWD needs zin");
    MC_CHECK(p->bias_mode == MC_BIAS_NONE || p->bias != nullptr, "gemm: bias_mode set without bias");
    SimtArgs g{};
    g.M = (int)p->M; g.N = (int)p->N; g.K = (int)p->K; g.batch = (int)p->batch;
    g.k_spans_batch = p->k_spans_batch ? 1 : 0;
    g.out_batch = g.k_spans_batch ? 1 : g.batch;
    g.A = reinterpret_cast<const float*>(p->A);
    g.B = reinterpret_cast<const float*>(p->B);
    g.a_kfast = p->a_major == MC_MAJOR_K; g.b_kfast = p->b_major == MC_MAJOR_K;
    g.a_sm = g.a_kfast ? p->lda : 1; g.a_sk = g.a_kfast ? 1 : p->lda; g.a_bs = p->a_batch_stride;
    g.b_sn = g.b_kfast ? p->ldb : 1; g.b_sk = g.b_kfast ? 1 : p->ldb; g.b_bs = p->b_batch_stride;
    g.C = reinterpret_cast<float*>(p->C); g.ldc = p->ldc; g.c_bs = p->c_batch_stride;
    g.accumulate = p->accumulate || p->split_k > 1; g.row_remap = p->row_remap;
    g.bias = p->bias; g.bias_mode = p->bias_mode;
    g.zout = reinterpret_cast<float*>(p->zout); g.ldz = p->ldz; g.z_bs = p->z_batch_stride;
    g.zin = reinterpret_cast<const float*>(p->zin); g.ldzin = p->ldzin; g.zin_bs = p->zin_batch_stride;
    g.act = p->act; g.R = p->R; g.ldr = p->ldr; g.r_bs = p->r_batch_stride;
    g.rowsum_out = p->rowsum_out;
    g.c_transposed = p->c_transposed ? 1 : 0;
    MC_CHECK(!p->c_transposed || (p->zout == nullptr && p->zin == nullptr && p->row_remap == 0),
             "gemm: c_transposed excludes zout / zin / row_remap");
    const long long tiles = ceil_div(p->M, TM) * ceil_div(p->N, TN);
    MC_CHECK(tiles < (1ll << 31) && g.out_batch < 65536, "simt gemm: grid too large");
    // The small projection GEMMs (model.py:288,424: 256 x 512 x 768) fill 32 tiles and walk K in 48 dependent
    // load -> sync -> FMA rounds: split K over blockIdx.z until the grid covers the machine (plain epilogue only).
    g.splits = 1;
    g.k_per_split = g.K;
    const bool plain = p->bias_mode == MC_BIAS_NONE && p->zout == nullptr && p->act == MC_ACT_NONE && p->R == nullptr &&
                       p->rowsum_out == nullptr && !g.k_spans_batch && g.out_batch == 1 && p->row_remap == 0 &&
                       !p->c_transposed && (g.accumulate || p->ldc == p->N);
    if (plain && tiles < 2 * sm_count()) {
        int sp = (int)ceil_div(2 * sm_count(), tiles);
        const int max_sp = (int)(g.K / 64);
        if (sp > max_sp) sp = max_sp;
        if (sp > 1) {
            g.k_per_split = (int)(ceil_div(ceil_div(g.K, sp), TK) * TK);
            g.splits = (int)ceil_div(g.K, g.k_per_split);
        }
    }
    if (g.splits > 1 && !g.accumulate) {
        MC_CUDA(cudaMemsetAsync(g.C, 0, sizeof(float) * (size_t)p->M * (size_t)p->N, stream));
    }
    dim3 grid((unsigned)tiles, (unsigned)g.out_batch, (unsigned)g.splits);
    MC_LAUNCH((gemm_simt_kernel), grid, 256, 0, stream, g);
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}
