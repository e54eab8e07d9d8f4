// Shared device/host helpers for the b200u kernels (sm_100a only).
//
// Everything in csrc/ is written for one target: -gencode arch=compute_100a,code=sm_100a.
// The library never allocates or frees device memory and never synchronises; every entry
// point enqueues on the caller's stream so the whole step can be captured in a CUDA graph.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// ---------------------------------------------------------------------------------------
// Error plumbing: C-ABI returns 0 / negative code, message via b200u_last_error_string().
// ---------------------------------------------------------------------------------------
namespace b200u {

void set_error(const char* fmt, ...);

#define B200U_OK 0
#define B200U_ERR_ARG -1
#define B200U_ERR_CUDA -2
#define B200U_ERR_UNSUPPORTED -3

#define B200U_CHECK_ARG(cond, ...)                                     \
    do {                                                               \
        if (!(cond)) {                                                 \
            ::b200u::set_error(__VA_ARGS__);                           \
            return B200U_ERR_ARG;                                      \
        }                                                              \
    } while (0)

void count_launch();  // bumps the kernel-launch counter read by b200u_launch_count()

#define B200U_CHECK_LAUNCH(name)                                                        \
    do {                                                                                \
        ::b200u::count_launch();                                                        \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess) {                                                        \
            ::b200u::set_error("%s: launch failed: %s", name, cudaGetErrorString(_e));  \
            return B200U_ERR_CUDA;                                                      \
        }                                                                               \
    } while (0)

#define B200U_CHECK_CUDA(expr)                                                               \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::b200u::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));              \
            return B200U_ERR_CUDA;                                                           \
        }                                                                                    \
    } while (0)

bool prof_begin(cudaStream_t st, double flops, int* slot);
void prof_end(cudaStream_t st, int slot);
int num_sms();  // cached cudaDevAttrMultiProcessorCount of the current device
int ensure_dyn_smem(const void* kern, size_t bytes);  // per-device, thread-safe MaxDynamicSharedMemorySize >= bytes
unsigned* dev_err_ptr();  // per-device input-error word (b200u_input_errors), bits below
enum { ERR_WORD_ID = 1, ERR_POS_ID = 2, ERR_TYPE_ID = 4, ERR_GATHER_INDEX = 8, ERR_SCATTER_ID = 16 };
bool pdl_enabled();  // b200u_set_pdl(): launch kernels with programmatic stream serialization

// Every kernel of the library is launched through launch_k(): with PDL on, the launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs may become resident (and run
// their prologue up to pdl_wait()) while the preceding kernel on the stream is still draining.
// Contract: every kernel calls pdl_wait() before its first global-memory access (read OR write)
// and before any early return, then pdl_launch() so its own successor can be scheduled.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                            cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// ---------------------------------------------------------------------------------------
// Counter-based dropout RNG.
// One call yields 32 bits = two 16-bit uniform samples for an element PAIR; an element is
// kept when its sample >= thresh16 (thresh16 = round(p * 65536)). The mapping depends only on
// (seed, stream, pair index), so forward and backward kernels regenerate identical masks
// without storing them. `seed` lives in device memory so a captured CUDA graph sees a fresh
// value on every replay.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rng_mix(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
// 64 well-mixed bits from a 32-bit index (two 16-bit samples per word): multiply / xor-shift rounds; every
// 16-bit half is uniform and the four halves are pairwise independent (checked numerically, see
// tests/test_gpu_parity.py::test_gemm_dropout_statistics and the attention dropout tests).
__device__ __forceinline__ void hash64(uint32_t idx_xor_key, uint32_t& hi, uint32_t& lo) {
    uint32_t x = idx_xor_key * 0x9E3779B1u;
    x ^= x >> 15;
    lo = x * 0x85EBCA77u;
    lo ^= lo >> 16;
    hi = (x ^ (x >> 13)) * 0xC2B2AE3Du;
    hi ^= hi >> 16;
}
// 64 hash bits serve FOUR consecutive elements (two pairs): key = f(seed, stream) is loop-invariant, the
// per-quad work is xor, multiply, xor-shift and one 32 x 32 -> 64 bit multiply; the pair's word is the upper
// (even pair) or lower (odd pair) half. Callers that walk consecutive pairs compute each quad's hash once
// (common sub-expression), i.e. ~1.5 integer instructions per element instead of ~10.
__device__ __forceinline__ uint32_t rng_pair(uint64_t seed, uint32_t stream, uint32_t pair_idx) {
    const uint32_t key = rng_mix((uint32_t)seed ^ (stream * 0x9E3779B9U)) ^ (uint32_t)(seed >> 32);
    uint32_t hi, lo;
    hash64((pair_idx >> 1) ^ key, hi, lo);
    return (pair_idx & 1u) ? lo : hi;
}
// the two words of quad `quad_idx`: hi = rng_pair(.., 2 * quad_idx), lo = rng_pair(.., 2 * quad_idx + 1)
__device__ __forceinline__ void rng_quad(uint64_t seed, uint32_t stream, uint32_t quad_idx, uint32_t& hi, uint32_t& lo) {
    const uint32_t key = rng_mix((uint32_t)seed ^ (stream * 0x9E3779B9U)) ^ (uint32_t)(seed >> 32);
    hash64(quad_idx ^ key, hi, lo);
}
// the four pair words of 8 consecutive elements starting at element index elem0 (a multiple of 8)
__device__ __forceinline__ void rng_words8(uint64_t seed, uint32_t stream, size_t elem0, uint32_t (&h)[4]) {
    const uint32_t q = (uint32_t)(elem0 >> 2);
    rng_quad(seed, stream, q, h[0], h[1]);
    rng_quad(seed, stream, q + 1, h[2], h[3]);
}
struct DropoutCfg {
    const unsigned long long* seed_ptr;  // device pointer, may be null when thresh16 == 0
    uint32_t stream;                     // distinct per (layer, site)
    uint32_t thresh16;                   // 0 => dropout disabled
    float scale;                         // 1 / (1 - p)
};
__device__ __forceinline__ uint64_t load_seed(const DropoutCfg& d) {
    return d.thresh16 ? *d.seed_ptr : 0ull;
}

// Dropout on the attention probabilities (model/layer.py:95). One hash serves a 2x2 block of the [L, L]
// probability matrix of a (sample, head): idx = (bh * L2 + (i >> 1)) * L2 + (j >> 1), L2 = (L + 1) >> 1,
//   (hi, lo) = hash64(idx ^ key)   (two multiply / xor-shift rounds, see hash64)
// query row i uses the word (i & 1) ? lo : hi, key column j its (j & 1) ? upper : lower 16 bits, and the
// probability is kept when that sample >= thresh16. A thread that walks a row (forward: thread = query)
// or a column (backward: thread = key) of the matrix therefore needs one hash per two scores either way,
// and every kernel (tcgen05 and mma.sync, forward and backward) regenerates the same mask.
__device__ __forceinline__ uint32_t attn_key(uint64_t seed, uint32_t stream) {
    return rng_mix((uint32_t)seed ^ (stream * 0x9E3779B9U)) ^ (uint32_t)(seed >> 32);
}
__device__ __forceinline__ uint32_t attn_block_base(int bh, int i, int L) {
    const int L2 = (L + 1) >> 1;
    return (uint32_t)((bh * L2 + (i >> 1)) * L2);
}
__device__ __forceinline__ void attn_rng_block(uint32_t key, uint32_t block_idx, uint32_t& hi, uint32_t& lo) {
    hash64(block_idx ^ key, hi, lo);
}
// the word of query-row parity `i_odd` of block `block_idx` (= attn_block_base(bh, i, L) + (j >> 1))
__device__ __forceinline__ uint32_t attn_rng(uint32_t key, uint32_t block_idx, int i_odd) {
    uint32_t hi, lo;
    attn_rng_block(key, block_idx, hi, lo);
    return i_odd ? lo : hi;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {   // device-wide nanosecond clock (bring-up stamps)
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Programmatic dependent launch (see launch_k above). No-ops when launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
    pdl_wait();
    pdl_launch();
}

// ---------------------------------------------------------------------------------------
// Small math helpers (reference: model/layer.py:31-37 exact erf GELU).
// ---------------------------------------------------------------------------------------
// erf via Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7, i.e. fp32-level for GELU): one
// reciprocal, one exp2 and a handful of FMAs instead of erff()'s ~40-instruction polynomial,
// evaluated on element PAIRS with packed fp32x2 instructions so the fused GEMM epilogues stay
// cheaper than the main loop they overlap with.
__device__ __forceinline__ float rcp_approx(float x) {  // MUFU.RCP, callers guarantee x >= 1
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {  // MUFU.EX2
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// ---- packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2(float lo, float hi) {
    f32x2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ f32x2 f2s(float c) { return f2(c, c); }
__device__ __forceinline__ void f2_get(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// two bf16 packed in a 32-bit word (low half = first element) <-> fp32 pair
__device__ __forceinline__ f32x2 f2_from_bf16x2(uint32_t u) {
    return f2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf16x2(f32x2 v) {
    float lo, hi;
    f2_get(v, lo, hi);
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// q = erfc(a * s / sqrt 2) for a >= 0 via Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7) and
// e = exp(-(a s)^2 / 2), on a PAIR of values: 2 MUFU.RCP + 2 MUFU.EX2 + 8 packed FMA/MUL.
// `a` is |x| pre-scaled by 1/s (the GELU forward passes |x|/2, so s = 2).
template <int S>
__device__ __forceinline__ f32x2 erfc_pair(f32x2 a, f32x2& e) {
    const f32x2 d = f2_fma(a, f2s(0.3275911f * 0.70710678118654752440f * S), f2s(1.0f));
    float d0, d1;
    f2_get(d, d0, d1);
    const f32x2 t = f2(rcp_approx(d0), rcp_approx(d1));
    f32x2 p = f2_fma(t, f2s(1.061405429f), f2s(-1.453152027f));
    p = f2_fma(p, t, f2s(1.421413741f));
    p = f2_fma(p, t, f2s(-0.284496736f));
    p = f2_fma(p, t, f2s(0.254829592f));
    const f32x2 s = f2_mul(f2_mul(a, a), f2s(-0.72134752044448170368f * S * S));
    float s0, s1;
    f2_get(s, s0, s1);
    e = f2(ex2_approx(s0), ex2_approx(s1));
    return f2_mul(f2_mul(p, t), e);
}
// gelu(x) = x/2 (1 + erf(x/sqrt 2)) = relu(x) - |x|/2 * erfc(|x|/sqrt 2)   (model/layer.py:31-37)
__device__ __forceinline__ f32x2 gelu_pair(f32x2 x) {
    const f32x2 h = f2_mul(x, f2s(0.5f));
    float h0, h1;
    f2_get(h, h0, h1);
    const f32x2 ah = f2(fabsf(h0), fabsf(h1));
    f32x2 e;
    const f32x2 q = erfc_pair<2>(ah, e);
    return f2_fma(f2_mul(ah, q), f2s(-1.0f), f2_add(h, ah));
}
// gelu'(x) = Phi(x) + x phi(x), Phi(x) = 1/2 + copysign(1/2 - erfc(|x|/sqrt 2)/2, x)
__device__ __forceinline__ f32x2 gelu_grad_pair(f32x2 x) {
    float x0, x1;
    f2_get(x, x0, x1);
    const f32x2 ax = f2(fabsf(x0), fabsf(x1));
    f32x2 e;
    const f32x2 q = erfc_pair<1>(ax, e);
    const f32x2 hq = f2_fma(q, f2s(-0.5f), f2s(0.5f));  // in [0, 1/2]
    float q0, q1;
    f2_get(hq, q0, q1);
    const f32x2 sg = f2(__uint_as_float(__float_as_uint(q0) | (__float_as_uint(x0) & 0x80000000u)),
                        __uint_as_float(__float_as_uint(q1) | (__float_as_uint(x1) & 0x80000000u)));
    const f32x2 cdf = f2_add(sg, f2s(0.5f));
    return f2_fma(f2_mul(x, e), f2s(0.39894228040143267794f), cdf);
}
// gelu'(x) (returned) and gelu(x) (val) from one erfc / exp evaluation: the FFN1 forward epilogue saves
// the derivative so the backward epilogue is a plain multiply.
__device__ __forceinline__ f32x2 gelu_both_pair(f32x2 x, f32x2& val) {
    float x0, x1;
    f2_get(x, x0, x1);
    const f32x2 ax = f2(fabsf(x0), fabsf(x1));
    f32x2 e;
    const f32x2 q = erfc_pair<1>(ax, e);                 // erfc(|x| / sqrt 2), e = exp(-x^2 / 2)
    const f32x2 hq = f2_fma(q, f2s(-0.5f), f2s(0.5f));   // 1/2 - q/2 in [0, 1/2]
    float q0, q1;
    f2_get(hq, q0, q1);
    const f32x2 sg = f2(__uint_as_float(__float_as_uint(q0) | (__float_as_uint(x0) & 0x80000000u)),
                        __uint_as_float(__float_as_uint(q1) | (__float_as_uint(x1) & 0x80000000u)));
    const f32x2 cdf = f2_add(sg, f2s(0.5f));             // Phi(x)
    // gelu(x) = relu(x) - |x|/2 * erfc(|x| / sqrt 2): no cancellation for negative x
    val = f2_fma(f2_mul(ax, q), f2s(-0.5f), f2_mul(f2_add(x, ax), f2s(0.5f)));
    return f2_fma(f2_mul(x, e), f2s(0.39894228040143267794f), cdf);
}
__device__ __forceinline__ float gelu_erf(float x) {
    float lo, hi;
    f2_get(gelu_pair(f2(x, x)), lo, hi);
    return lo;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
    float lo, hi;
    f2_get(gelu_grad_pair(f2(x, x)), lo, hi);
    return lo;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    bf162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    bf162 v = *reinterpret_cast<bf162*>(&u);
    return __bfloat1622float2(v);
}

// ---------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, tcgen05 / TMEM.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
// non-blocking probe (try_wait may suspend the thread for a hardware time-out before it answers false)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 2-D tiled TMA load: box lands at `dst` (swizzled as the tensor map says), completion is
// signalled as transaction bytes on `bar`. c0 = innermost (contiguous) coordinate.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// Same with the destination and barrier given as shared-space addresses (kept in uniform registers
// by the warp-uniform producer loop of the GEMM).
__device__ __forceinline__ void tma_load_2d_u(uint32_t dst, const CUtensorMap* tmap, uint32_t bar,
                                              int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_u(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// smem (swizzled box) -> global, bulk-group completion; OOB parts of the box are clipped.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
// global += smem box (element-wise add in the tensor map's data type, performed at L2).
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tmap, const void* src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_but_one() {
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Multicast variant: the box lands at the same smem offset in every CTA of `cta_mask` and
// completes transaction bytes on the mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* tmap, uint64_t* bar,
                                               int c0, int c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
// tcgen05.commit that arrives on the barrier at this offset in every CTA of `cta_mask`.
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
// true in exactly one (converged) lane of the warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// ---- distributed shared memory: address of the same smem location in CTA `rank` of the cluster, remote
// stores, remote mbarrier arrival (release at cluster scope) and the matching acquire-wait
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void dsmem_st_f32x2(uint32_t raddr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(raddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void dsmem_mbar_arrive(uint32_t rbar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
// remote 8-byte store that completes 8 transaction bytes on the (remote) mbarrier `rbar` when it lands:
// the data and its arrival signal travel together, no fence / separate arrive needed
__device__ __forceinline__ void dsmem_st_async_f32x2(uint32_t raddr, float a, float b, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                 ::"r"(raddr), "f"(a), "f"(b), "r"(rbar) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS)
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on `bar` once every tcgen05.mma previously issued by this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// TMEM -> registers: this thread's lane (row), 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp semantics restated).
// ---------------------------------------------------------------------------------------
// smem matrix descriptor, SWIZZLE_128B, version 1 (Blackwell).
//  K-major : rows of 128 B (64 k-elements); 8-row groups SBO = 1024 B apart; LBO unused.
//  MN-major: rows of 128 B (64 mn-elements), one row per k; 8-k groups SBO = 1024 B apart;
//            next 64-wide mn block LBO bytes away (= one TMA box = BLOCK_K * 128 B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
// descriptor without the start address (constant per operand layout)
__host__ __device__ constexpr uint64_t desc_template(uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor for kind::f16, bf16 x bf16 -> f32.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn) {
    return (1u << 4)                      // D format f32
           | (1u << 7)                    // A bf16
           | (1u << 10)                   // B bf16
           | ((a_mn ? 1u : 0u) << 15)     // A major
           | ((b_mn ? 1u : 0u) << 16)     // B major
           | ((uint32_t)(n >> 3) << 17)   // N
           | ((uint32_t)(m >> 4) << 24);  // M
}

// 2-D bf16 / f32 tensor map of a row-major [rows, cols] matrix (leading dimension ld elements): boxes of
// box_rows x 128 bytes, 128B-swizzled. Host side, defined in runtime.cu.
int make_tmap(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld, int box_rows, bool f32 = false);

int make_tmap_3d(CUtensorMap* tm, const void* ptr, int batch, int rows, int cols, int ld, int box_rows);
// 3-D tiled TMA store (smem box -> global, clipped per dimension), bulk-group completion.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tmap, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// Vector fp32 reduction into global memory (split-K wgrad accumulation).
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}

}  // namespace b200u
