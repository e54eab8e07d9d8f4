// K7 — pooler, small classification heads and the pos-weighted BCE loss (sm_100a).
//
//  pooler      : tanh(h[:, 0] · Wᵀ + b)                        model/layer.py:179-185
//  small linear: x · Wᵀ + b for a handful of classes            model/meme_uniter.py:20, pretrain.py:62
//  bce_logits  : BCEWithLogitsLoss(pos_weight) mean + d/dlogit  train_template.py:64-65,98-99
// These are latency kernels (B <= a few hundred rows): fp32 master weights are read directly,
// everything accumulates in fp32, and every launch is a single small grid.
#include "../../include/b200u.h"
#include "common.cuh"

namespace b200u {

// pooled[b, n] = tanh(b[n] + sum_k W[n,k] * h[b*row_stride + k]);  h is bf16.
__global__ void __launch_bounds__(256)
pooler_fwd_kernel(const bf16* __restrict__ h, size_t row_stride, const float* __restrict__ W,
                  const float* __restrict__ bias, float* __restrict__ pooled, int H) {
    pdl_sync();
    extern __shared__ float sx[];  // [H]
    const int b = blockIdx.y;
    for (int k = threadIdx.x; k < H; k += blockDim.x) sx[k] = __bfloat162float(h[(size_t)b * row_stride + k]);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = 0; i < 8; ++i) {
        const int n = blockIdx.x * 64 + warp * 8 + i;
        if (n >= H) break;
        const float* w = W + (size_t)n * H;
        float acc = 0.f;
        for (int k = lane * 4; k < H; k += 128) {
            float4 wv = *reinterpret_cast<const float4*>(w + k);
            acc += wv.x * sx[k] + wv.y * sx[k + 1] + wv.z * sx[k + 2] + wv.w * sx[k + 3];
        }
        acc = warp_sum(acc);
        if (lane == 0) pooled[(size_t)b * H + n] = tanhf(acc + bias[n]);
    }
}

// blocks [0, ceil(H/8)):  dW[n,:] += sum_b dpre[b,n] h0[b,:] ; db[n] += sum_b dpre[b,n]
// blocks [ceil(H/8), +B*ceil(H/64)): dh0[b, 64-slice] = sum_n dpre[b,n] W[n, slice]   (written as bf16)
// with dpre = dpooled * (1 - pooled^2).
__global__ void __launch_bounds__(256)
pooler_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ pooled,
                  const bf16* __restrict__ h, size_t row_stride, const float* __restrict__ W,
                  float* __restrict__ dW, float* __restrict__ db, bf16* __restrict__ dh,
                  size_t dh_row_stride, int B, int H) {
    pdl_sync();
    extern __shared__ float sm[];
    const int nwb = (H + 7) / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if ((int)blockIdx.x < nwb) {
        // one warp per output row n of dW: lane b keeps dpre[b, n] (32 samples at a time) and the
        // sample loop broadcasts it with a shuffle, so the h rows stream without dependent loads
        const int n = blockIdx.x * 8 + warp;
        if (n >= H) return;
        float bsum = 0.f;
        for (int b0 = 0; b0 < B; b0 += 32) {
            const int bl = b0 + lane;
            float mine = 0.f;
            if (bl < B) {
                const float p = pooled[(size_t)bl * H + n];
                mine = dpooled[(size_t)bl * H + n] * (1.0f - p * p);
            }
            bsum += mine;
            const int nb = min(32, B - b0);
            for (int kk = 0; kk < H; kk += 128) {  // warp-uniform trip count (the shuffles need all lanes)
                const int k0 = kk + lane * 4;
                const bool on = k0 < H;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
                for (int b = 0; b < nb; ++b) {
                    const float dpre = __shfl_sync(0xffffffffu, mine, b);
                    if (on) {
                        uint2 u = *reinterpret_cast<const uint2*>(h + (size_t)(b0 + b) * row_stride + k0);
                        float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y);
                        a0 = fmaf(dpre, f0.x, a0); a1 = fmaf(dpre, f0.y, a1);
                        a2 = fmaf(dpre, f1.x, a2); a3 = fmaf(dpre, f1.y, a3);
                    }
                }
                if (on) {
                    float4* dst = reinterpret_cast<float4*>(dW + (size_t)n * H + k0);
                    float4 cur = *dst;
                    cur.x += a0; cur.y += a1; cur.z += a2; cur.w += a3;
                    *dst = cur;
                }
            }
        }
        bsum = warp_sum(bsum);
        if (lane == 0 && db) db[n] += bsum;
    } else {
        // one block per (sample, 64-column slice): 4 groups of 64 threads stride over n, then reduce
        const int nkb = (H + 63) / 64;
        const int idx = blockIdx.x - nwb;
        const int b = idx / nkb, k0 = (idx - b * nkb) * 64;
        float* sd = sm;       // dpre[b, :]
        float* red = sm + H;  // [4][64]
        for (int n = threadIdx.x; n < H; n += blockDim.x) {
            const float p = pooled[(size_t)b * H + n];
            sd[n] = dpooled[(size_t)b * H + n] * (1.0f - p * p);
        }
        __syncthreads();
        const int kc = threadIdx.x & 63, ng = threadIdx.x >> 6;
        const int k = k0 + kc;
        float acc = 0.f;
        if (k < H) {
#pragma unroll 8
            for (int n = ng; n < H; n += 4) acc = fmaf(sd[n], W[(size_t)n * H + k], acc);
        }
        red[ng * 64 + kc] = acc;
        __syncthreads();
        if (ng == 0 && k < H)
            dh[(size_t)b * dh_row_stride + k] = __float2bfloat16(red[kc] + red[64 + kc] + red[128 + kc] + red[192 + kc]);
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

// blocks [0, B): dx[b,:] = sum_c dout[b,c] W[c,:] ; blocks [B, B+C): dW[c,:] += sum_b dout[b,c] x[b,:],
// db[c] += sum_b dout[b,c].
__global__ void __launch_bounds__(256)
linear_small_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                        const float* __restrict__ W, float* __restrict__ dx, float* __restrict__ dW,
                        float* __restrict__ db, int B, int C, int K) {
    pdl_sync();
    if ((int)blockIdx.x < B) {
        const int b = blockIdx.x;
        if (!dx) return;
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            float acc = 0.f;
            for (int c = 0; c < C; ++c) acc = fmaf(dout[(size_t)b * C + c], W[(size_t)c * K + k], acc);
            dx[(size_t)b * K + k] = acc;
        }
    } else {
        const int c = blockIdx.x - B;
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            float acc = 0.f;
            for (int b = 0; b < B; ++b) acc = fmaf(dout[(size_t)b * C + c], x[(size_t)b * K + k], acc);
            dW[(size_t)c * K + k] += acc;
        }
        if (threadIdx.x == 0 && db) {
            float s = 0.f;
            for (int b = 0; b < B; ++b) s += dout[(size_t)b * C + c];
            db[c] += s;
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
    B200U_CHECK_ARG(h && W && bias && pooled && H % 4 == 0, "pooler_fwd: bad arguments");
    if (B == 0) return B200U_OK;
    launch_k(pooler_fwd_kernel, dim3(dim3((H + 63) / 64, B)), dim3(256), H * sizeof(float), stream, (const bf16*)h, (size_t)row_stride, W, bias, pooled, H);
    B200U_CHECK_LAUNCH("pooler_fwd");
    return B200U_OK;
}

extern "C" int b200u_pooler_bwd(const float* dpooled, const float* pooled, const void* h,
                                long long row_stride, const float* W, float* dW, float* db, void* dh,
                                long long dh_row_stride, int B, int H, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(dpooled && pooled && h && W && dW && dh && H % 4 == 0 && row_stride % 4 == 0, "pooler_bwd: bad arguments");
    if (B == 0) return B200U_OK;
    launch_k(pooler_bwd_kernel, dim3((H + 7) / 8 + B * ((H + 63) / 64)), dim3(256), (H + 256) * sizeof(float), stream, dpooled, pooled, (const bf16*)h, (size_t)row_stride, W, dW, db, (bf16*)dh, (size_t)dh_row_stride, B, H);
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
    launch_k(linear_small_bwd_kernel, dim3(B + C), dim3(256), 0, stream, dout, x, W, dx, dW, db, B, C, K);
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
