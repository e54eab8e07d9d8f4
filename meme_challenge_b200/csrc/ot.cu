// K8 — optimal-transport word/region alignment (reference model/ot.py), fp32 throughout.
//
//   cosine_cost      : 1 - normalize(x)·normalize(y)ᵀ, joint padding zeroed      ot.py:11-21,72-75
//   ipot             : the whole 50-iteration IPOT loop in ONE launch              ot.py:35-66
//   ot_distance      : trace(C · T) = sum_ij C[i,j] T[j,i]                         ot.py:24-32,84
//   cosine_cost_bwd  : gradient of the distance through the cost (T is detached)  ot.py:82-84
//
// The reference issues ~7 tiny kernels per IPOT iteration (~350 launches per call, pure launch
// latency). Here one CTA owns one sample: A = exp(-Cᵀ/beta) and T live in shared memory for all
// iterations, the two mat-vecs per iteration are warp-shuffle reductions, and nothing but the final
// plan is written to HBM.
#include "../../include/b200u.h"
#include "common.cuh"

namespace b200u {

// inv[r] = 1 / max(||v_r||, eps)   (F.normalize semantics), one warp per row.
__global__ void __launch_bounds__(256)
row_inv_norm_kernel(const float* __restrict__ v, float* __restrict__ inv, int rows, int D, float eps) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float a = v[(size_t)r * D + d];
        s = fmaf(a, a, s);
    }
    s = warp_sum(s);
    if (lane == 0) inv[r] = 1.0f / fmaxf(sqrtf(s), eps);
}

// cost[b,m,n] = pad ? 0 : 1 - <x[b,m], y[b,n]> * xinv * yinv. grid (N tiles of 8 warps, M, B).
__global__ void __launch_bounds__(256)
cosine_cost_kernel(const float* __restrict__ x, const float* __restrict__ y,
                   const float* __restrict__ xinv, const float* __restrict__ yinv,
                   const unsigned char* __restrict__ x_pad, const unsigned char* __restrict__ y_pad,
                   float* __restrict__ cost, int M, int N, int D) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int m = blockIdx.y, b = blockIdx.z;
    if (n >= N) return;
    const float* xr = x + ((size_t)b * M + m) * D;
    const float* yr = y + ((size_t)b * N + n) * D;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(xr[d], yr[d], s);
    s = warp_sum(s);
    if (lane == 0) {
        const bool pad = (x_pad && x_pad[(size_t)b * M + m]) || (y_pad && y_pad[(size_t)b * N + n]);
        cost[((size_t)b * M + m) * N + n] = pad ? 0.f : 1.0f - s * xinv[(size_t)b * M + m] * yinv[(size_t)b * N + n];
    }
}

// One CTA per sample. smem: A[N][Mp], Q/T[N][Mp], sigma[M], delta[N].
__global__ void __launch_bounds__(256)
ipot_kernel(const float* __restrict__ C, const unsigned char* __restrict__ x_pad,
            const unsigned char* __restrict__ y_pad, float* __restrict__ T_out, int M, int N, int Mp,
            float beta, int iterations, int k_inner) {
    pdl_sync();
    extern __shared__ float sm[];
    float* A = sm;
    float* T = A + (size_t)N * Mp;
    float* sigma = T + (size_t)N * Mp;
    float* delta = sigma + M;
    __shared__ float s_len[2];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const unsigned char* xp = x_pad + (size_t)b * M;
    const unsigned char* yp = y_pad + (size_t)b * N;

    // lengths = number of non-padded positions (ot.py:77-80)
    if (warp == 0) {
        int c = 0;
        for (int m = lane; m < M; m += 32) c += xp[m] ? 0 : 1;
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) s_len[0] = (float)c;
    } else if (warp == 1) {
        int c = 0;
        for (int n = lane; n < N; n += 32) c += yp[n] ? 0 : 1;
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) s_len[1] = (float)c;
    }
    __syncthreads();
    const float x_len = s_len[0], y_len = s_len[1];
    // sigma = 1/x_len (0 on pad); T = 1, A = exp(-Cᵀ/beta) (both 0 on joint pad)   ot.py:38-47
    for (int m = tid; m < M; m += blockDim.x) sigma[m] = xp[m] ? 0.f : 1.0f / x_len;
    for (int i = tid; i < N * M; i += blockDim.x) {
        const int n = i / M, m = i - n * M;
        const bool pad = xp[m] || yp[n];
        A[n * Mp + m] = pad ? 0.f : expf(-C[((size_t)b * M + m) * N + n] / beta);
        T[n * Mp + m] = pad ? 0.f : 1.0f;
    }
    __syncthreads();

    for (int it = 0; it < iterations; ++it) {
        // Q = A * T (kept in T's storage)                                            ot.py:58
        for (int i = tid; i < N * M; i += blockDim.x) {
            const int n = i / M, m = i - n * M;
            T[n * Mp + m] *= A[n * Mp + m];
        }
        __syncthreads();
        for (int kk = 0; kk < k_inner; ++kk) {
            // delta = 1 / (y_len * Q·sigma + 1e4*y_pad)                              ot.py:62
            for (int n = warp; n < N; n += nw) {
                float s = 0.f;
                for (int m = lane; m < M; m += 32) s = fmaf(T[n * Mp + m], sigma[m], s);
                s = warp_sum(s);
                if (lane == 0) delta[n] = 1.0f / (y_len * s + (yp[n] ? 1e4f : 0.f));
            }
            __syncthreads();
            // sigma = 1 / (x_len * delta·Q + 1e4*x_pad)                              ot.py:63
            for (int m = warp; m < M; m += nw) {
                float s = 0.f;
                for (int n = lane; n < N; n += 32) s = fmaf(delta[n], T[n * Mp + m], s);
                s = warp_sum(s);
                if (lane == 0) sigma[m] = 1.0f / (x_len * s + (xp[m] ? 1e4f : 0.f));
            }
            __syncthreads();
        }
        // T = delta * Q * sigma                                                      ot.py:64
        for (int i = tid; i < N * M; i += blockDim.x) {
            const int n = i / M, m = i - n * M;
            T[n * Mp + m] = delta[n] * T[n * Mp + m] * sigma[m];
        }
        __syncthreads();
    }
    // final mask + store [B, N, M]                                                   ot.py:65-66
    for (int i = tid; i < N * M; i += blockDim.x) {
        const int n = i / M, m = i - n * M;
        T_out[(size_t)b * N * M + i] = (xp[m] || yp[n]) ? 0.f : T[n * Mp + m];
    }
}

// dist[b] = sum_{m,n} C[b,m,n] * T[b,n,m]
__global__ void __launch_bounds__(256)
ot_distance_kernel(const float* __restrict__ C, const float* __restrict__ T, float* __restrict__ dist,
                   int M, int N) {
    pdl_sync();
    __shared__ float red[8];
    const int b = blockIdx.x;
    float s = 0.f;
    for (int i = threadIdx.x; i < M * N; i += blockDim.x) {
        const int m = i / N, n = i - m * N;
        s = fmaf(C[(size_t)b * M * N + i], T[((size_t)b * N + n) * M + m], s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        dist[b] = t;
    }
}

// Backward of dist through the cost: dC[m,n] = ddist * T[n,m] (0 on joint pad), then through
// cost = 1 - xn·ynᵀ and xn = x * xinv:  dxn = -dC·yn ; dx = (dxn - xn <xn, dxn>) * xinv.
// which == 0: rows of x (grid.x = M), which == 1: rows of y (grid.x = N). One CTA per row.
__global__ void __launch_bounds__(256)
cosine_cost_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                       const float* __restrict__ xinv, const float* __restrict__ yinv,
                       const unsigned char* __restrict__ x_pad, const unsigned char* __restrict__ y_pad,
                       const float* __restrict__ T, const float* __restrict__ ddist,
                       float* __restrict__ dx, float* __restrict__ dy, int M, int N, int D) {
    pdl_sync();
    extern __shared__ float sm[];  // coefficients of the other side's rows
    __shared__ float red[8];
    const int b = blockIdx.z, which = blockIdx.y, r = blockIdx.x;
    const int rows = which == 0 ? M : N, other = which == 0 ? N : M;
    if (r >= rows) return;
    const float gd = ddist[b];
    const unsigned char* pr = which == 0 ? x_pad : y_pad;
    const unsigned char* po = which == 0 ? y_pad : x_pad;
    const float* self = (which == 0 ? x + ((size_t)b * M + r) * D : y + ((size_t)b * N + r) * D);
    const float sinv = which == 0 ? xinv[(size_t)b * M + r] : yinv[(size_t)b * N + r];
    const float* oth = which == 0 ? y + (size_t)b * N * D : x + (size_t)b * M * D;
    const float* oinv = which == 0 ? yinv + (size_t)b * N : xinv + (size_t)b * M;
    float* out = which == 0 ? dx + ((size_t)b * M + r) * D : dy + ((size_t)b * N + r) * D;
    const bool rpad = pr && pr[(size_t)b * rows + r];
    // coef[o] = -dC[r,o] * oinv[o]   (dC = ddist * T[n,m], zero on padding)
    for (int o = threadIdx.x; o < other; o += blockDim.x) {
        const int m = which == 0 ? r : o, n = which == 0 ? o : r;
        const bool pad = rpad || (po && po[(size_t)b * other + o]);
        sm[o] = pad ? 0.f : -gd * T[((size_t)b * N + n) * M + m] * oinv[o];
    }
    __syncthreads();
    // dxn[d] = sum_o coef[o] * oth[o, d] ; then project out the radial component
    float dot = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float a = 0.f;
        for (int o = 0; o < other; ++o) a = fmaf(sm[o], oth[(size_t)o * D + d], a);
        out[d] = a;  // stash dxn
        dot = fmaf(a, self[d] * sinv, dot);
    }
    dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    for (int d = threadIdx.x; d < D; d += blockDim.x) out[d] = (out[d] - self[d] * sinv * tot) * sinv;
}

}  // namespace b200u

using namespace b200u;

extern "C" int b200u_cosine_cost(const float* x, const float* y, const unsigned char* x_pad,
                                 const unsigned char* y_pad, float* cost, float* xinv, float* yinv,
                                 int B, int M, int N, int D, float eps, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(x && y && cost && xinv && yinv && M > 0 && N > 0 && D > 0, "cosine_cost: bad arguments");
    if (B == 0) return B200U_OK;
    launch_k(row_inv_norm_kernel, dim3((B * M + 7) / 8), dim3(256), 0, stream, x, xinv, B * M, D, eps);
    B200U_CHECK_LAUNCH("row_inv_norm(x)");
    launch_k(row_inv_norm_kernel, dim3((B * N + 7) / 8), dim3(256), 0, stream, y, yinv, B * N, D, eps);
    B200U_CHECK_LAUNCH("row_inv_norm(y)");
    launch_k(cosine_cost_kernel, dim3(dim3((N + 7) / 8, M, B)), dim3(256), 0, stream, x, y, xinv, yinv, x_pad, y_pad, cost, M, N, D);
    B200U_CHECK_LAUNCH("cosine_cost");
    return B200U_OK;
}

extern "C" int b200u_ipot(const float* cost, const unsigned char* x_pad, const unsigned char* y_pad,
                          float* T, int B, int M, int N, float beta, int iterations, int k,
                          b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(cost && x_pad && y_pad && T && M > 0 && N > 0 && iterations >= 0 && k >= 1, "ipot: bad arguments");
    if (B == 0) return B200U_OK;
    const int Mp = M | 1;  // odd row stride: conflict-free column walks
    const size_t smem = ((size_t)2 * N * Mp + M + N) * sizeof(float);
    B200U_CHECK_ARG(smem <= 220 * 1024, "ipot: %d x %d plan does not fit in shared memory", N, M);
    if (int rc = ensure_dyn_smem((const void*)ipot_kernel, smem)) return rc;
    launch_k(ipot_kernel, dim3(B), dim3(256), smem, stream, cost, x_pad, y_pad, T, M, N, Mp, beta, iterations, k);
    B200U_CHECK_LAUNCH("ipot");
    return B200U_OK;
}

extern "C" int b200u_ot_distance(const float* cost, const float* T, float* dist, int B, int M, int N,
                                 b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(cost && T && dist, "ot_distance: null pointer");
    if (B == 0) return B200U_OK;
    launch_k(ot_distance_kernel, dim3(B), dim3(256), 0, stream, cost, T, dist, M, N);
    B200U_CHECK_LAUNCH("ot_distance");
    return B200U_OK;
}

extern "C" int b200u_cosine_cost_bwd(const float* x, const float* y, const float* xinv, const float* yinv,
                                     const unsigned char* x_pad, const unsigned char* y_pad, const float* T,
                                     const float* ddist, float* dx, float* dy, int B, int M, int N, int D,
                                     b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(x && y && xinv && yinv && T && ddist && dx && dy, "cosine_cost_bwd: null pointer");
    if (B == 0) return B200U_OK;
    const int mx = M > N ? M : N;
    launch_k(cosine_cost_bwd_kernel, dim3(dim3(mx, 2, B)), dim3(256), mx * sizeof(float), stream, x, y, xinv, yinv, x_pad, y_pad, T, ddist, dx, dy, M, N, D);
    B200U_CHECK_LAUNCH("cosine_cost_bwd");
    return B200U_OK;
}
