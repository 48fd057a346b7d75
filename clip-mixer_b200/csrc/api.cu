// Library-level entry points of the C ABI: version, thread-local error message, device info.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace mc {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

// Number of SMs the persistent kernels size their grids for.  MC_SM_LIMIT (environment, or mc_set_sm_limit) leaves
// some SMs to concurrent work: in data-parallel runs the NCCL all-reduce of the gradient buckets needs a few CTAs
// resident WHILE the backward GEMMs run, and a persistent grid that fills every SM (each CTA takes a whole SM's shared
// memory) would push the collective into the gaps between kernels.
// Two separate values: the environment default is read once and survives; mc_set_sm_limit(n > 0) overrides it
// (FusedTrainStep sizes each tower's kernels for its share of the SMs) and mc_set_sm_limit(0) RESTORES the environment
// default instead of discarding it.
static int g_env_limit = -1;
static int g_sm_override = 0;

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;  // B200
    }
    if (g_env_limit < 0) {
        const char* v = getenv("MC_SM_LIMIT");
        g_env_limit = v ? atoi(v) : 0;
        if (g_env_limit < 0) g_env_limit = 0;
    }
    int n = cached;
    int lim = g_sm_override > 0 ? g_sm_override : g_env_limit;
    if (g_sm_override > 0 && g_env_limit > 0 && g_env_limit < lim) lim = g_env_limit;
    if (lim > 0 && lim < n) n = lim;
    return n & ~1;   // CTA pairs: keep it even
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MC_PDL");
        v = e ? (atoi(e) != 0) : 0;
    }
    return v != 0;
}

}  // namespace mc

extern "C" int mc_version(void) { return 100; }

extern "C" int mc_set_sm_limit(int sms) {
    mc::g_sm_override = sms > 0 ? sms : 0;
    return MC_OK;
}

extern "C" const char* mc_last_error(void) { return mc::g_last_error; }

extern "C" int mc_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    MC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MC_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return MC_OK;
}
