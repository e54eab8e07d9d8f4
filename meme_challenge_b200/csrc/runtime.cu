// Error reporting and device queries for the b200u C-ABI.
#include "../../include/b200u.h"
#include "common.cuh"

#include <stdarg.h>

namespace b200u {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

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
    return cached;
}

}  // namespace b200u

extern "C" const char* b200u_last_error_string(void) { return b200u::g_err; }

extern "C" int b200u_version(void) { return 100; }

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
