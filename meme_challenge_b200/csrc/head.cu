// K7 — pooler, small classification heads and the pos-weighted BCE loss (sm_100a).
//
//  pooler      : tanh(h[:, 0] · Wᵀ + b)                        model/layer.py:179-185
//  small linear: x · Wᵀ + b for a handful of classes            model/meme_uniter.py:20, pretrain.py:62
//  bce_logits  : BCEWithLogitsLoss(pos_weight) mean + d/dlogit  train_template.py:64-65,98-99
// These are latency kernels (B <= a few hundred rows): fp32 master weights are read directly,
// everything accumulates in fp32, and every launch is a single small grid.
#include "../../include/b200u.h"
#include "common.cuh"

#include <algorithm>
#include <mutex>

namespace b200u {

// v[0..31] per lane -> lane l returns the sum over the warp's lanes of v[l] (31 shuffles)
__device__ __forceinline__ float transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const bool up = (lane & step) != 0;
#pragma unroll
        for (int k = 0; k < step; ++k) {
            const float send = up ? v[k] : v[k + step];
            const float recv = __shfl_xor_sync(0xffffffffu, send, step);
            v[k] = (up ? v[k + step] : v[k]) + recv;
        }
    }
    return v[0];
}

constexpr int POOL_MAXV = 8;   // float4 per lane: H <= 1024
constexpr int POOL_BC = 32;    // samples per chunk (= lanes)

// first-token rows of samples [b0, b0 + nb) -> fp32 rows in shared memory; 16-byte loads, 6 in flight per thread
__device__ __forceinline__ void pool_fill_rows(float* sx, const bf16* __restrict__ h, size_t row_stride, int b0,
                                               int nb, int H) {
    const int hv = H >> 3, nvec = nb * hv;
    for (int i0 = 0; i0 < nvec; i0 += 6 * 256) {
        uint4 u[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int i = i0 + j * 256 + threadIdx.x;
            if (i < nvec) {
                const int bb = i / hv, v = i - bb * hv;
                u[j] = *reinterpret_cast<const uint4*>(h + (size_t)(b0 + bb) * row_stride + v * 8);
            }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int i = i0 + j * 256 + threadIdx.x;
            if (i < nvec) {
                const int bb = i / hv, v = i - bb * hv;
                const float2 f0 = unpack_bf16(u[j].x), f1 = unpack_bf16(u[j].y), f2 = unpack_bf16(u[j].z), f3 = unpack_bf16(u[j].w);
                float4* dst = reinterpret_cast<float4*>(sx + (size_t)bb * H + v * 8);
                dst[0] = make_float4(f0.x, f0.y, f1.x, f1.y);
                dst[1] = make_float4(f2.x, f2.y, f3.x, f3.y);
            }
        }
    }
}

// pooled[b, n] = tanh(bias[n] + sum_k W[n,k] * h[b*row_stride + k]);  h is bf16.
// One warp per output feature n (its weight row stays in registers), samples in chunks of 32 whose first-token
// rows sit in shared memory as fp32: every weight is read from global exactly once per CTA instead of once per
// sample, and the 32 dot products of a chunk are reduced with one 31-shuffle transpose.
__global__ void __launch_bounds__(256)
pooler_fwd_kernel(const bf16* __restrict__ h, size_t row_stride, const float* __restrict__ W,
                  const float* __restrict__ bias, float* __restrict__ pooled, int B, int H) {
    pdl_sync();
    extern __shared__ __align__(16) float sx[];  // [32][H]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * 8 + warp;
    const int nv = H >> 7;  // float4 per lane
    float4 w[POOL_MAXV];
#pragma unroll
    for (int i = 0; i < POOL_MAXV; ++i)
        w[i] = (i < nv && n < H) ? *reinterpret_cast<const float4*>(W + (size_t)n * H + (i * 32 + lane) * 4) : make_float4(0, 0, 0, 0);
    const float bn = n < H ? bias[n] : 0.f;
    for (int b0 = 0; b0 < B; b0 += POOL_BC) {
        const int nb = min(POOL_BC, B - b0);
        __syncthreads();
        pool_fill_rows(sx, h, row_stride, b0, nb, H);
        __syncthreads();
        float acc[32];
#pragma unroll
        for (int bb = 0; bb < 32; ++bb) {
            float a = 0.f;
            if (bb < nb) {
#pragma unroll
                for (int i = 0; i < POOL_MAXV; ++i) {
                    if (i < nv) {
                        const float4 x = *reinterpret_cast<const float4*>(sx + (size_t)bb * H + (i * 32 + lane) * 4);
                        a = fmaf(w[i].x, x.x, fmaf(w[i].y, x.y, fmaf(w[i].z, x.z, fmaf(w[i].w, x.w, a))));
                    }
                }
            }
            acc[bb] = a;
        }
        const float tot = transpose_sum32(acc, lane);   // lane l: sample b0 + l
        if (lane < nb && n < H) pooled[(size_t)(b0 + lane) * H + n] = tanhf(tot + bn);
    }
}

// blocks [0, ceil(H/8)): one warp per output feature n: dW[n,:] += sum_b dpre[b,n] h0[b,:] ; db[n] += sum_b dpre[b,n]
//   (h0 rows of a 32-sample chunk in shared memory, dpre[b, n] broadcast by shuffle, one plain read-modify-write
//    of the warp's own dW row)
// blocks [ceil(H/8), +H/KS): dh0[b, KS-column slice] = sum_n dpre[b,n] W[n, slice] (written as bf16). The CTA pulls its
//   W column slice ([H][KS] fp32, 32-byte row segments, every load in flight at once) and the chunk's dpre
//   (sample-minor, [H][32]) into shared memory; a thread owns a 4-sample x 4-column block of the output for one
//   share of n (two 16-byte shared loads per 16 FMAs) and the n shares meet in shared memory.
// with dpre = dpooled * (1 - pooled^2).
template <int KS>
__global__ void __launch_bounds__(256)
pooler_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ pooled,
                  const bf16* __restrict__ h, size_t row_stride, const float* __restrict__ W,
                  float* __restrict__ dW, float* __restrict__ db, bf16* __restrict__ dh,
                  size_t dh_row_stride, int B, int H) {
    pdl_sync();
    extern __shared__ __align__(16) float sm[];
    const int nkb = H / KS;   // the dh0 blocks come first: they are the longer ones
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if ((int)blockIdx.x >= nkb) {
        float* sx = sm;  // [32][H] first-token rows of the chunk
        const int n = (blockIdx.x - nkb) * 8 + warp;
        const int nv = H >> 7;
        float4 acc[POOL_MAXV];
#pragma unroll
        for (int i = 0; i < POOL_MAXV; ++i) acc[i] = make_float4(0, 0, 0, 0);
        float bsum = 0.f;
        for (int b0 = 0; b0 < B; b0 += POOL_BC) {
            const int nb = min(POOL_BC, B - b0);
            __syncthreads();
            pool_fill_rows(sx, h, row_stride, b0, nb, H);
            float mine = 0.f;
            if (lane < nb && n < H) {
                const float p = pooled[(size_t)(b0 + lane) * H + n];
                mine = dpooled[(size_t)(b0 + lane) * H + n] * (1.0f - p * p);
            }
            __syncthreads();
            bsum += mine;
#pragma unroll 4
            for (int bb = 0; bb < nb; ++bb) {
                const float dpre = __shfl_sync(0xffffffffu, mine, bb);
#pragma unroll
                for (int i = 0; i < POOL_MAXV; ++i) {
                    if (i < nv) {
                        const float4 x = *reinterpret_cast<const float4*>(sx + (size_t)bb * H + (i * 32 + lane) * 4);
                        acc[i].x = fmaf(dpre, x.x, acc[i].x); acc[i].y = fmaf(dpre, x.y, acc[i].y);
                        acc[i].z = fmaf(dpre, x.z, acc[i].z); acc[i].w = fmaf(dpre, x.w, acc[i].w);
                    }
                }
            }
        }
        if (n < H) {
#pragma unroll
            for (int i = 0; i < POOL_MAXV; ++i) {
                if (i < nv) {
                    float4* dst = reinterpret_cast<float4*>(dW + (size_t)n * H + (i * 32 + lane) * 4);
                    float4 cur = *dst;
                    cur.x += acc[i].x; cur.y += acc[i].y; cur.z += acc[i].z; cur.w += acc[i].w;
                    *dst = cur;
                }
            }
            bsum = warp_sum(bsum);
            if (lane == 0 && db) db[n] += bsum;
        }
    } else {
        constexpr int TPG = 8 * (KS / 4);      // threads per n share: 8 sample quads x KS/4 column quads
        constexpr int NG = 256 / TPG;          // n shares
        float* sW = sm;                        // [H][KS]
        float* sd = sm + (size_t)H * KS;       // [H][32] dpre, sample-minor
        float* red = sd + (size_t)H * 32;      // [NG][32][KS]
        const int k0 = blockIdx.x * KS;
        {   // W[:, k0 .. k0+KS): one 16-byte load per (row, column quad)
            constexpr int QPR = KS / 4;
            const int nvec = H * QPR;
            for (int i0 = 0; i0 < nvec; i0 += 8 * 256) {
                float4 u[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int i = i0 + j * 256 + threadIdx.x;
                    if (i < nvec) u[j] = *reinterpret_cast<const float4*>(W + (size_t)(i / QPR) * H + k0 + (i % QPR) * 4);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int i = i0 + j * 256 + threadIdx.x;
                    if (i < nvec) *reinterpret_cast<float4*>(sW + (size_t)i * 4) = u[j];
                }
            }
        }
        const int g = threadIdx.x / TPG, t = threadIdx.x % TPG;
        const int bq = t & 7, kq = t >> 3;
        const int n_per = (H + NG - 1) / NG;
        for (int b0 = 0; b0 < B; b0 += POOL_BC) {
            const int nb = min(POOL_BC, B - b0);
            __syncthreads();
            // dpre of the chunk: lane = sample (conflict-free transposed stores), 16-byte loads along n
            const int nq = H >> 2;
            for (int q0 = warp; q0 < nq; q0 += 8 * 8) {
                float4 dp[8], pp[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int q = q0 + j * 8;
                    if (q < nq && lane < nb) {
                        dp[j] = *reinterpret_cast<const float4*>(dpooled + (size_t)(b0 + lane) * H + 4 * q);
                        pp[j] = *reinterpret_cast<const float4*>(pooled + (size_t)(b0 + lane) * H + 4 * q);
                    } else {
                        dp[j] = pp[j] = make_float4(0, 0, 0, 0);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int q = q0 + j * 8;
                    if (q < nq) {
                        sd[(4 * q + 0) * 32 + lane] = dp[j].x * (1.0f - pp[j].x * pp[j].x);
                        sd[(4 * q + 1) * 32 + lane] = dp[j].y * (1.0f - pp[j].y * pp[j].y);
                        sd[(4 * q + 2) * 32 + lane] = dp[j].z * (1.0f - pp[j].z * pp[j].z);
                        sd[(4 * q + 3) * 32 + lane] = dp[j].w * (1.0f - pp[j].w * pp[j].w);
                    }
                }
            }
            __syncthreads();
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            const int n_lo = g * n_per, n_hi = min(H, n_lo + n_per);
#pragma unroll 4
            for (int n = n_lo; n < n_hi; ++n) {
                const float4 d = *reinterpret_cast<const float4*>(sd + (size_t)n * 32 + 4 * bq);
                const float4 wv = *reinterpret_cast<const float4*>(sW + (size_t)n * KS + 4 * kq);
                const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] = fmaf(dv[i], wv.x, acc[i][0]); acc[i][1] = fmaf(dv[i], wv.y, acc[i][1]);
                    acc[i][2] = fmaf(dv[i], wv.z, acc[i][2]); acc[i][3] = fmaf(dv[i], wv.w, acc[i][3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(red + ((size_t)g * 32 + 4 * bq + i) * KS + 4 * kq) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            __syncthreads();
            for (int i = threadIdx.x; i < 32 * KS; i += 256) {
                const int bb = i / KS, kk = i - bb * KS;
                float s_ = 0.f;
#pragma unroll
                for (int gg = 0; gg < NG; ++gg) s_ += red[((size_t)gg * 32 + bb) * KS + kk];
                if (bb < nb) dh[(size_t)(b0 + bb) * dh_row_stride + k0 + kk] = __float2bfloat16(s_);
            }
        }
    }
}

// out[b, c] = bias[c] + sum_k x[b,k] W[c,k]; one warp per (b, c).
__global__ void __launch_bounds__(256)
linear_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                        const float* __restrict__ bias, float* __restrict__ out, int B, int C, int K) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (idx >= B * C) return;
    const int b = idx / C, c = idx - b * C;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(x[(size_t)b * K + k], W[(size_t)c * K + k], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[idx] = acc + (bias ? bias[c] : 0.f);
}

// blocks [0, nxb): one thread per (sample, 4 columns): dx[b, k..k+3] = sum_c dout[b,c] W[c, k..k+3]
// blocks [nxb, +C * ceil(K/256)): thread per (class c, column k): dW[c,k] += sum_b dout[b,c] x[b,k] with 16 row
// loads in flight; the first block of each class also does db[c] += sum_b dout[b,c].
__global__ void __launch_bounds__(256)
linear_small_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                        const float* __restrict__ W, float* __restrict__ dx, float* __restrict__ dW,
                        float* __restrict__ db, int B, int C, int K, int nxb) {
    pdl_sync();
    if ((int)blockIdx.x < nxb) {
        const int kq = K >> 2;
        const int idx = blockIdx.x * 256 + threadIdx.x;
        if (!dx || idx >= B * kq) return;
        const int b = idx / kq, q = idx - b * kq;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
            const float dv = dout[(size_t)b * C + c];
            const float4 w = *reinterpret_cast<const float4*>(W + (size_t)c * K + 4 * q);
            acc.x = fmaf(dv, w.x, acc.x); acc.y = fmaf(dv, w.y, acc.y);
            acc.z = fmaf(dv, w.z, acc.z); acc.w = fmaf(dv, w.w, acc.w);
        }
        *reinterpret_cast<float4*>(dx + (size_t)b * K + 4 * q) = acc;
    } else {
        const int kb = (K + 255) / 256;
        const int c = (blockIdx.x - nxb) / kb, k = ((blockIdx.x - nxb) % kb) * 256 + threadIdx.x;
        if (k < K) {
            float acc = 0.f;
            for (int b0 = 0; b0 < B; b0 += 16) {
                float xv[16], dv[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const bool ok = b0 + u < B;
                    xv[u] = ok ? x[(size_t)(b0 + u) * K + k] : 0.f;
                    dv[u] = ok ? dout[(size_t)(b0 + u) * C + c] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) acc = fmaf(dv[u], xv[u], acc);
            }
            dW[(size_t)c * K + k] += acc;
        }
        if (db && (blockIdx.x - nxb) % kb == 0 && threadIdx.x < 32) {
            float s_ = 0.f;
            for (int b = threadIdx.x; b < B; b += 32) s_ += dout[(size_t)b * C + c];
            s_ = warp_sum(s_);
            if (threadIdx.x == 0) db[c] += s_;
        }
    }
}

// loss = mean_b[(1-y) x + (1 + (w-1) y) softplus(-x)] ; dlogit = [(1-y) - (1+(w-1)y) sigmoid(-x)] * gscale / B
// (the numerically stable form torch.nn.BCEWithLogitsLoss uses). Single block.
__global__ void __launch_bounds__(256)
bce_logits_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                  float pos_weight, float grad_scale, float* __restrict__ loss, float* __restrict__ dlogits,
                  float* __restrict__ probs, int B) {
    pdl_sync();
    __shared__ float red[8];
    float part = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float x = logits[b], y = labels[b];
        const float c = 1.0f + (pos_weight - 1.0f) * y;
        const float sp = log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f);  // softplus(-x)
        part += (1.0f - y) * x + c * sp;
        const float sig_neg = 1.0f / (1.0f + expf(x));  // sigmoid(-x)
        if (dlogits) dlogits[b] = ((1.0f - y) - c * sig_neg) * grad_scale / (float)B;
        if (probs) probs[b] = 1.0f / (1.0f + expf(-x));
    }
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        if (loss) *loss = s / (float)B;
    }
}

}  // namespace b200u

using namespace b200u;

extern "C" int b200u_pooler_fwd(const void* h, long long row_stride, const float* W, const float* bias,
                                float* pooled, int B, int H, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(h && W && bias && pooled && H % 128 == 0 && H <= 1024 && row_stride % 8 == 0,
                    "pooler_fwd: bad arguments (H must be a multiple of 128, <= 1024)");
    if (B == 0) return B200U_OK;
    const size_t smem = (size_t)POOL_BC * H * sizeof(float);
    static std::mutex mu;
    static size_t set_for[64] = {};
    int dev = 0;
    B200U_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64) {
        std::lock_guard<std::mutex> lock(mu);
        if (smem > set_for[dev]) {
            B200U_CHECK_CUDA(cudaFuncSetAttribute(pooler_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set_for[dev] = smem;
        }
    }
    launch_k(pooler_fwd_kernel, dim3((H + 7) / 8), dim3(256), smem, stream, (const bf16*)h, (size_t)row_stride, W, bias, pooled, B, H);
    B200U_CHECK_LAUNCH("pooler_fwd");
    return B200U_OK;
}

extern "C" int b200u_pooler_bwd(const float* dpooled, const float* pooled, const void* h,
                                long long row_stride, const float* W, float* dW, float* db, void* dh,
                                long long dh_row_stride, int B, int H, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(dpooled && pooled && h && W && dW && dh && H % 128 == 0 && H <= 1024 && row_stride % 8 == 0,
                    "pooler_bwd: bad arguments (H must be a multiple of 128, <= 1024)");
    if (B == 0) return B200U_OK;
    // dh0 blocks own 8-column slices of W (H/8 of them, as many as the dW blocks): [H][8] W slice + [H][32] dpre + partials
    constexpr int KS = 8;
    const size_t smem = std::max((size_t)POOL_BC * H, (size_t)H * KS + (size_t)H * 32 + (size_t)(256 / (8 * (KS / 4))) * 32 * KS) * sizeof(float);
    static std::mutex mu;
    static size_t set_for[64] = {};
    int dev = 0;
    B200U_CHECK_CUDA(cudaGetDevice(&dev));
    auto kern = pooler_bwd_kernel<KS>;
    if (dev >= 0 && dev < 64) {
        std::lock_guard<std::mutex> lock(mu);
        if (smem > set_for[dev]) {
            B200U_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set_for[dev] = smem;
        }
    }
    launch_k(kern, dim3((H + 7) / 8 + H / KS), dim3(256), smem, stream, dpooled, pooled, (const bf16*)h, (size_t)row_stride, W, dW, db, (bf16*)dh, (size_t)dh_row_stride, B, H);
    B200U_CHECK_LAUNCH("pooler_bwd");
    return B200U_OK;
}

extern "C" int b200u_linear_small_fwd(const float* x, const float* W, const float* bias, float* out,
                                      int B, int C, int K, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(x && W && out, "linear_small_fwd: null pointer");
    if (B * C == 0) return B200U_OK;
    launch_k(linear_small_fwd_kernel, dim3((B * C + 7) / 8), dim3(256), 0, stream, x, W, bias, out, B, C, K);
    B200U_CHECK_LAUNCH("linear_small_fwd");
    return B200U_OK;
}

extern "C" int b200u_linear_small_bwd(const float* dout, const float* x, const float* W, float* dx,
                                      float* dW, float* db, int B, int C, int K, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(dout && x && W && dW, "linear_small_bwd: null pointer");
    if (B * C == 0) return B200U_OK;
    B200U_CHECK_ARG(K % 4 == 0, "linear_small_bwd: K must be a multiple of 4");
    const int nxb = dx ? (B * (K / 4) + 255) / 256 : 0;
    launch_k(linear_small_bwd_kernel, dim3(nxb + C * ((K + 255) / 256)), dim3(256), 0, stream, dout, x, W, dx, dW, db, B, C, K, nxb);
    B200U_CHECK_LAUNCH("linear_small_bwd");
    return B200U_OK;
}

extern "C" int b200u_bce_logits(const float* logits, const float* labels, float pos_weight,
                                float grad_scale, float* loss, float* dlogits, float* probs, int B,
                                b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(logits && labels && B > 0, "bce_logits: bad arguments");
    launch_k(bce_logits_kernel, dim3(1), dim3(256), 0, stream, logits, labels, pos_weight, grad_scale, loss, dlogits, probs, B);
    B200U_CHECK_LAUNCH("bce_logits");
    return B200U_OK;
}
