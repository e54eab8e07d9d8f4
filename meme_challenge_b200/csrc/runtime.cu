// Error reporting and device queries for the b200u C-ABI.
#include "../../include/b200u.h"
#include "common.cuh"

#include <map>
#include <mutex>
#include <utility>

#include <cudaTypedefs.h>
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

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// matrix stored as [rows, cols] row-major with leading dimension ld (elements); box = box_cols x
// box_rows with a 128-byte swizzled inner dimension (64 bf16 or 32 f32 columns).
int make_tmap(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld, int box_rows, bool f32) {
    auto fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return B200U_ERR_CUDA;
    }
    const int esz = f32 ? 4 : 2;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                    const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%d cols=%d ld=%d box_rows=%d f32=%d",
                  (int)r, ptr, rows, cols, ld, box_rows, (int)f32);
        return B200U_ERR_CUDA;
    }
    return B200U_OK;
}

// [batch, rows, cols] bf16 view of a row-major [batch * rows, cols] matrix (leading dimension ld): boxes of
// 1 x box_rows x 64 columns, 128B-swizzled. Rows past `rows` are out of bounds PER SAMPLE: loads zero-fill
// them, stores clip them, so a tile that overhangs a sample never touches its neighbour.
int make_tmap_3d(CUtensorMap* tm, const void* ptr, int batch, int rows, int cols, int ld, int box_rows) {
    auto fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return B200U_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * ld * 2};
    cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d) failed (%d) ptr=%p batch=%d rows=%d cols=%d ld=%d box_rows=%d", (int)r,
                  ptr, batch, rows, cols, ld, box_rows);
        return B200U_ERR_CUDA;
    }
    return B200U_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device and PyTorch calls the library from its main and autograd
// threads: remember, per (device, kernel), the largest size already granted, under one mutex.
static std::mutex g_smem_mu;
static std::map<std::pair<int, const void*>, size_t> g_smem_set;
int ensure_dyn_smem(const void* kern, size_t bytes) {
    int dev = 0;
    B200U_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_smem_mu);
    size_t& have = g_smem_set[std::make_pair(dev, kern)];
    if (bytes > have) {
        B200U_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
    return B200U_OK;
}

// Per-device 4-byte input-error word (the one allocation the library owns): kernels OR a bit into it when an index they
// were handed is out of range, and substitute a safe value instead of reading or writing out of bounds.
static std::mutex g_flag_mu;
static unsigned* g_flag[64] = {};
unsigned* dev_err_ptr() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_flag_mu);
    if (!g_flag[dev]) {
        unsigned* p = nullptr;
        if (cudaMalloc(&p, sizeof(unsigned)) != cudaSuccess) return nullptr;
        cudaMemset(p, 0, sizeof(unsigned));
        g_flag[dev] = p;
    }
    return g_flag[dev];
}

}  // namespace b200u

extern "C" int b200u_input_errors(unsigned* bits, int reset) {
    using namespace b200u;
    B200U_CHECK_ARG(bits, "input_errors: null pointer");
    unsigned* p = dev_err_ptr();
    B200U_CHECK_ARG(p, "input_errors: no error word on this device");
    B200U_CHECK_CUDA(cudaDeviceSynchronize());
    B200U_CHECK_CUDA(cudaMemcpy(bits, p, sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (reset) B200U_CHECK_CUDA(cudaMemset(p, 0, sizeof(unsigned)));
    return B200U_OK;
}

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
