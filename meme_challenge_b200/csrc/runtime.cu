// Error reporting and device queries for the b200u C-ABI.
#include "../../include/b200u.h"
#include "common.cuh"

#include <stdarg.h>
#include <stdlib.h>

namespace b200u {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }

// ---- optional per-GEMM event timing (bench.py roofline leg; never on inside graph capture) ----
struct ProfRec { cudaEvent_t e0, e1; double flops; };
static ProfRec* g_prof = nullptr;
static int g_prof_cap = 0, g_prof_n = 0;
static bool g_prof_on = false;
bool prof_begin(cudaStream_t st, double flops, int* slot) {
    if (!g_prof_on || g_prof_n >= g_prof_cap) return false;
    *slot = g_prof_n++;
    g_prof[*slot].flops = flops;
    cudaEventRecord(g_prof[*slot].e0, st);
    return true;
}
void prof_end(cudaStream_t st, int slot) { cudaEventRecord(g_prof[slot].e1, st); }

static int g_pdl = 1;
bool pdl_enabled() { return g_pdl != 0; }
void set_pdl(int on) { g_pdl = on; }

static int g_sm_limit = 0;  // b200u_set_sm_limit(): persistent grids are sized to at most this many SMs
void set_sm_limit(int n) { g_sm_limit = n; }

int num_sms() {
    static int cached = 0;
    if (!cached) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return (g_sm_limit > 0 && g_sm_limit < cached) ? g_sm_limit : cached;
}

}  // namespace b200u

extern "C" const char* b200u_last_error_string(void) { return b200u::g_err; }

extern "C" int b200u_version(void) { return 100; }

extern "C" int b200u_set_sm_limit(int n) {
    b200u::set_sm_limit(n);
    return B200U_OK;
}

extern "C" int b200u_set_pdl(int on) {
    b200u::set_pdl(on);
    return B200U_OK;
}

extern "C" long long b200u_launch_count(void) { return __atomic_load_n(&b200u::g_launches, __ATOMIC_RELAXED); }

extern "C" int b200u_prof_enable(int max_records) {
    using namespace b200u;
    if (max_records > g_prof_cap) {
        ProfRec* n = (ProfRec*)realloc(g_prof, sizeof(ProfRec) * (size_t)max_records);
        if (!n) { set_error("prof_enable: out of host memory"); return B200U_ERR_ARG; }
        g_prof = n;
        for (int i = g_prof_cap; i < max_records; ++i) {
            B200U_CHECK_CUDA(cudaEventCreate(&g_prof[i].e0));
            B200U_CHECK_CUDA(cudaEventCreate(&g_prof[i].e1));
        }
        g_prof_cap = max_records;
    }
    g_prof_n = 0;
    g_prof_on = max_records > 0;
    return B200U_OK;
}

extern "C" int b200u_prof_collect(double* total_ms, double* total_flops, int* count) {
    using namespace b200u;
    g_prof_on = false;
    double ms = 0.0, fl = 0.0;
    for (int i = 0; i < g_prof_n; ++i) {
        B200U_CHECK_CUDA(cudaEventSynchronize(g_prof[i].e1));
        float t = 0.f;
        B200U_CHECK_CUDA(cudaEventElapsedTime(&t, g_prof[i].e0, g_prof[i].e1));
        ms += t;
        fl += g_prof[i].flops;
    }
    if (total_ms) *total_ms = ms;
    if (total_flops) *total_flops = fl;
    if (count) *count = g_prof_n;
    g_prof_n = 0;
    return B200U_OK;
}

extern "C" int b200u_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    B200U_CHECK_CUDA(cudaGetDevice(&dev));
    int n = 0, ma = 0, mi = 0;
    B200U_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    B200U_CHECK_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
    B200U_CHECK_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = ma;
    if (cc_minor) *cc_minor = mi;
    return B200U_OK;
}
