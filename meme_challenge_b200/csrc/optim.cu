// K10 — fused optimizer step over the flat parameter / gradient buffers (sm_100a, HBM-bound).
//
// Replaces, for the whole model in three launches, the per-tensor sequence of the reference:
//   average_gradients (grad /= accum)            train_template.py:89-92
//   torch.nn.utils.clip_grad_norm_(params, 5)    train_template.py:104
//   torch.optim.Adam(L2 weight decay groups)     utils/optim_utils.py:16-46, train_template.py:105
// and additionally refreshes the bf16 shadow copy of the weights that the tcgen05 GEMMs read.
// 16 B/param read + 14 B/param written -> 30 B/param per step.
#include "../../include/b200u.h"
#include "common.cuh"

namespace b200u {

// Gradient source of element i: the fp32 buffer g, except on [lo16, hi16) (multiples of 4) where the
// data-parallel step left the all-reduced gradients as bf16 in g16[i - lo16] (half the NVLink volume).
__device__ __forceinline__ float4 load_grad4(const float* __restrict__ g, const bf16* __restrict__ g16, size_t lo16,
                                             size_t hi16, size_t i0) {
    if (g16 && i0 >= lo16 && i0 < hi16) {
        const uint2 u = *reinterpret_cast<const uint2*>(g16 + (i0 - lo16));
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
        return make_float4(a.x, a.y, b.x, b.y);
    }
    return *reinterpret_cast<const float4*>(g + i0);
}
__device__ __forceinline__ float load_grad1(const float* __restrict__ g, const bf16* __restrict__ g16, size_t lo16,
                                            size_t hi16, size_t i) {
    return (g16 && i >= lo16 && i < hi16) ? __bfloat162float(g16[i - lo16]) : g[i];
}

// sumsq += sum g[i]^2  (double accumulation across blocks)
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, size_t n, double* __restrict__ sumsq, const bf16* __restrict__ g16,
             size_t lo16, size_t hi16) {
    pdl_sync();
    float acc = 0.f;
    const size_t nvec = n >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = load_grad4(g, g16, lo16, hi16, i << 2);
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (size_t i = nvec << 2; i < n; ++i) {
            const float x = load_grad1(g, g16, lo16, hi16, i);
            acc += x * x;
        }
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += (double)red[w];
        atomicAdd(sumsq, s);
    }
}

// coef = pre_scale * min(1, max_norm / (pre_scale * sqrt(sumsq) + 1e-6)); norm_out = pre_scale*sqrt(sumsq)
// (clip_grad_norm_ semantics applied to the already-averaged gradients)
__global__ void clip_coef_kernel(const double* __restrict__ sumsq, float pre_scale, float max_norm,
                                 float* __restrict__ coef, float* __restrict__ norm_out) {
    pdl_sync();
    const float norm = pre_scale * (float)sqrt(*sumsq);
    float c = 1.0f;
    if (max_norm > 0.f) {
        c = max_norm / (norm + 1e-6f);
        if (c > 1.0f) c = 1.0f;
    }
    *coef = pre_scale * c;
    if (norm_out) *norm_out = norm;
}

// Adam with L2 weight decay folded into the gradient (torch.optim.Adam, not AdamW):
//   g = coef*g + wd*p ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2
//   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// `runs` partitions [0,n) into segments of constant weight decay: run r covers
// [run_start[r], run_start[r+1]) with decay run_wd[r]; chunk_run[c] is the run containing the
// first element of chunk c (chunk = 1024 elements), so a thread only walks forward.
// run_wd[r] < 0 marks a run the optimizer must not touch (parameters that receive no gradient:
// torch.optim.Adam skips tensors whose .grad is None, e.g. mask_embedding during fine-tuning, and
// frozen parameters): weights, moments and shadow stay as they are, only the gradient is zeroed.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, bf16* __restrict__ shadow, size_t n,
            const long long* __restrict__ run_start, const float* __restrict__ run_wd,
            const int* __restrict__ chunk_run, int num_runs, const float* __restrict__ coef_ptr,
            const float* __restrict__ lr_ptr, const unsigned long long* __restrict__ step_ptr,
            float beta1, float beta2, float eps, int zero_grad, float* __restrict__ g_mut,
            const bf16* __restrict__ g16, size_t lo16, size_t hi16) {
    pdl_sync();
    const float coef = coef_ptr ? *coef_ptr : 1.0f;
    const float lr = *lr_ptr;
    // step index lives on the device so a captured CUDA graph advances it on every replay
    const float t = (float)(*step_ptr);
    const float bc1 = 1.0f - powf(beta1, t);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, t));
    const float step_size = lr / bc1;
    const size_t nchunks = (n + 1023) >> 10;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const size_t i0 = (c << 10) + threadIdx.x * 4;
        if (i0 >= n) continue;
        int r = chunk_run[c];
        while (r + 1 < num_runs && (long long)i0 >= run_start[r + 1]) ++r;
        if (i0 + 4 <= n && (r + 1 >= num_runs || (long long)(i0 + 4) <= run_start[r + 1])) {
            const float wd = run_wd[r];
            if (wd < 0.f) {
                if (zero_grad) *reinterpret_cast<float4*>(g_mut + i0) = make_float4(0.f, 0.f, 0.f, 0.f);
                continue;
            }
            float4 pv = *reinterpret_cast<float4*>(p + i0);
            float4 gv = load_grad4(g, g16, lo16, hi16, i0);
            float4 mv = *reinterpret_cast<float4*>(m + i0);
            float4 vv = *reinterpret_cast<float4*>(v + i0);
            float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gg = coef * gp[k] + wd * pp[k];
                mp[k] = beta1 * mp[k] + (1.0f - beta1) * gg;
                vp[k] = beta2 * vp[k] + (1.0f - beta2) * gg * gg;
                pp[k] -= step_size * mp[k] / (sqrtf(vp[k]) / bc2_sqrt + eps);
            }
            *reinterpret_cast<float4*>(p + i0) = pv;
            *reinterpret_cast<float4*>(m + i0) = mv;
            *reinterpret_cast<float4*>(v + i0) = vv;
            if (shadow) {
                uint2 o;
                o.x = pack_bf16(pp[0], pp[1]);
                o.y = pack_bf16(pp[2], pp[3]);
                *reinterpret_cast<uint2*>(shadow + i0) = o;
            }
            if (zero_grad) *reinterpret_cast<float4*>(g_mut + i0) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (size_t i = i0; i < i0 + 4 && i < n; ++i) {
                while (r + 1 < num_runs && (long long)i >= run_start[r + 1]) ++r;
                const float wd = run_wd[r];
                if (wd < 0.f) {
                    if (zero_grad) g_mut[i] = 0.f;
                    continue;
                }
                const float gg = coef * load_grad1(g, g16, lo16, hi16, i) + wd * p[i];
                const float mm = beta1 * m[i] + (1.0f - beta1) * gg;
                const float vv = beta2 * v[i] + (1.0f - beta2) * gg * gg;
                const float pn = p[i] - step_size * mm / (sqrtf(vv) / bc2_sqrt + eps);
                m[i] = mm; v[i] = vv; p[i] = pn;
                if (shadow) shadow[i] = __float2bfloat16(pn);
                if (zero_grad) g_mut[i] = 0.f;
            }
        }
    }
}

__global__ void counter_add_kernel(unsigned long long* c, unsigned long long inc) {
    pdl_sync(); *c += inc; }

}  // namespace b200u

using namespace b200u;

// *counter += inc on the stream (dropout seed / optimizer step counters; graph-capturable).
extern "C" int b200u_counter_add(unsigned long long* counter, unsigned long long inc,
                                 b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(counter, "counter_add: null pointer");
    launch_k(counter_add_kernel, dim3(1), dim3(1), 0, stream, counter, inc);
    B200U_CHECK_LAUNCH("counter_add");
    return B200U_OK;
}

extern "C" int b200u_grad_sumsq(const float* g, size_t n, double* sumsq, const void* g16, size_t lo16,
                                size_t hi16, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(g && sumsq && ((uintptr_t)g & 15) == 0, "grad_sumsq: bad arguments");
    B200U_CHECK_ARG(!g16 || (lo16 % 4 == 0 && hi16 % 4 == 0 && lo16 <= hi16 && hi16 <= n && ((uintptr_t)g16 & 7) == 0),
                    "grad_sumsq: bf16 gradient range must be 4-element aligned and inside [0, n)");
    if (n == 0) return B200U_OK;
    size_t grid = ((n >> 2) + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    launch_k(sumsq_kernel, dim3((int)grid), dim3(256), 0, stream, g, n, sumsq, (const bf16*)g16, lo16, hi16);
    B200U_CHECK_LAUNCH("grad_sumsq");
    return B200U_OK;
}

extern "C" int b200u_clip_coef(const double* sumsq, float pre_scale, float max_norm, float* coef,
                               float* norm_out, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(sumsq && coef, "clip_coef: null pointer");
    launch_k(clip_coef_kernel, dim3(1), dim3(1), 0, stream, sumsq, pre_scale, max_norm, coef, norm_out);
    B200U_CHECK_LAUNCH("clip_coef");
    return B200U_OK;
}

extern "C" int b200u_adam_step(float* p, float* g, float* m, float* v, void* shadow_bf16, size_t n,
                               const long long* run_start, const float* run_wd, const int* chunk_run,
                               int num_runs, const float* coef, const float* lr,
                               const unsigned long long* step, float beta1, float beta2, float eps,
                               int zero_grad, const void* g16, size_t lo16, size_t hi16, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(p && g && m && v && run_start && run_wd && chunk_run && lr && num_runs > 0 && step, "adam_step: bad arguments");
    B200U_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam_step: buffers must be 16-byte aligned");
    B200U_CHECK_ARG(!g16 || (lo16 % 4 == 0 && hi16 % 4 == 0 && lo16 <= hi16 && hi16 <= n && ((uintptr_t)g16 & 7) == 0),
                    "adam_step: bf16 gradient range must be 4-element aligned and inside [0, n)");
    if (n == 0) return B200U_OK;
    size_t nchunks = (n + 1023) >> 10;
    size_t grid = nchunks;
    const size_t cap = (size_t)num_sms() * 16;
    if (grid > cap) grid = cap;
    launch_k(adam_kernel, dim3((int)grid), dim3(256), 0, stream, p, g, m, v, (bf16*)shadow_bf16, n, run_start, run_wd, chunk_run, num_runs, coef, lr, step, beta1, beta2, eps, zero_grad, g, (const bf16*)g16, lo16, hi16);
    B200U_CHECK_LAUNCH("adam_step");
    return B200U_OK;
}
