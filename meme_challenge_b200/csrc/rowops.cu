// HBM/L2-bound row kernels of the UNITER path (sm_100a): LayerNorm fwd/bwd with fused
// dropout-mask and bias-grad, column sums, fp32->bf16 casts, the bit-exact gather_index
// concat (K1) and its backward. One warp owns one row; every access is a 16-byte vector.
//
// Reference arithmetic:
//   LayerNorm  : apex FusedLayerNorm(H, eps=1e-12) == (x-mean)/sqrt(var+eps)*gamma+beta with
//                biased variance (model/model.py:229,252,253,258; model/layer.py:108,149)
//   gather     : torch.gather(torch.cat([txt_emb, img_emb], 1), 1, gather_index) (model/model.py:330-333)
#include "../../include/b200u.h"
#include "common.cuh"

namespace b200u {

constexpr int LN_MAXV = 4;  // vectors (of 8 elements) per lane -> H <= 1024

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float (&f)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 t;
    t = unpack_bf16(u.x); f[0] = t.x; f[1] = t.y;
    t = unpack_bf16(u.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16(u.z); f[4] = t.x; f[5] = t.y;
    t = unpack_bf16(u.w); f[6] = t.x; f[7] = t.y;
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&f)[8]);
template <>
__device__ __forceinline__ void store8<bf16>(bf16* p, const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
    o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = o;
}
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// Row statistics over values held in registers (two-pass: mean, then centred variance).
__device__ __forceinline__ void row_stats(const float (&x)[LN_MAXV][8], int nv, int lane, int H,
                                          float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
        if (lane + 32 * i < nv)
#pragma unroll
            for (int j = 0; j < 8; ++j) s += x[i][j];
    mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
        if (lane + 32 * i < nv)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float d = x[i][j] - mean;
                q += d * d;
            }
    rstd = rsqrtf(warp_sum(q) / (float)H + eps);
}

// ---------------------------------------------------------------------------------------
// LayerNorm forward: y = (x - mean) * rstd * gamma + beta ; optional dropout on y.
// ---------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, TOut* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int M, int H, float eps, DropoutCfg drop) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nv = H >> 3;
    float v[LN_MAXV][8];
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
        if (lane + 32 * i < nv) load8(x + (size_t)row * H + (lane + 32 * i) * 8, v[i]);
    float mean, rstd;
    row_stats(v, nv, lane, H, eps, mean, rstd);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
    const uint64_t seed = load_seed(drop);
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nv) {
            float g[8], b[8], o[8];
            load8(gamma + vi * 8, g);
            load8(beta + vi * 8, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
            if (drop.thresh16) {
                uint32_t hw[4];
                rng_words8(seed, drop.stream, (size_t)row * H + vi * 8, hw);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t h = hw[j];
                    o[2 * j] = ((h & 0xffffu) >= drop.thresh16) ? o[2 * j] * drop.scale : 0.f;
                    o[2 * j + 1] = ((h >> 16) >= drop.thresh16) ? o[2 * j + 1] * drop.scale : 0.f;
                }
            }
            store8(y + (size_t)row * H + vi * 8, o);
        }
    }
}

// ---------------------------------------------------------------------------------------
// LayerNorm backward for y = LN(x):
//   dx = rstd * (g*dy - mean_H(g*dy) - xhat * mean_H(g*dy*xhat))
//   dgamma += sum_rows dy*xhat ; dbeta += sum_rows dy
// Fused extras for the BertSelfOutput / BertOutput pattern x = dropout(dense) + residual
// (model/layer.py:111-115,152-156): dz = dropout_mask(dx) (the dense output's grad, bf16) and
// dbias += sum_rows dz. dx itself is the residual-branch grad.
//
// Layout: 20 warps per CTA, one CTA per SM, each warp owns ONE row per pass (2624 rows / 148 SMs
// = 18 rows per CTA at the C2 shape, so a single pass): all of a row's loads are in flight at once
// and nothing is carried in registers between rows. The per-row column contributions (dy*xhat, dy,
// bf16(dz)) are staged in shared memory and summed down the columns by the same CTA, two columns
// per thread, so the global fp32 accumulation costs 3*H atomics per CTA and no cross-warp
// register reduction.
// ---------------------------------------------------------------------------------------
constexpr int LNB_WARPS = 20;
constexpr int LNB_THREADS = LNB_WARPS * 32;

template <typename TX, int NVMAX>  // NVMAX: 16-byte vectors per lane (3 covers H <= 768, 4 covers H <= 1024)
__global__ void __launch_bounds__(LNB_THREADS, 1)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const TX* __restrict__ x,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, bf16* __restrict__ dx, bf16* __restrict__ dz,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                     int M, int H, int rows_per_cta, int rows_per_pass, DropoutCfg drop,
                     int drop_on_input) {
    pdl_sync();
    extern __shared__ __align__(16) float stage[];  // [3][rows_per_pass][H]: dy*xhat | dy | bf16(dz)
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nv = H >> 3;
    const uint64_t seed = load_seed(drop);
    float* sG = stage;
    float* sB = stage + (size_t)rows_per_pass * H;
    float* sZ = stage + (size_t)2 * rows_per_pass * H;

    const int r_begin = blockIdx.x * rows_per_cta;
    const int r_end = min(M, r_begin + rows_per_cta);
    const int c0 = 2 * threadIdx.x;  // this thread's column pair in the column pass
    float cg0 = 0.f, cg1 = 0.f, cb0 = 0.f, cb1 = 0.f, cz0 = 0.f, cz1 = 0.f;

    for (int base = r_begin; base < r_end; base += rows_per_pass) {
        const int row = base + warp;
        if (warp < rows_per_pass && row < r_end) {
            const float mu = mean[row], rs = rstd[row];
            float xh[NVMAX][8], gd[NVMAX][8];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < NVMAX; ++i) {
                const int vi = lane + 32 * i;
                if (vi < nv) {
                    load8(x + (size_t)row * H + vi * 8, xh[i]);
                    load8(dy + (size_t)row * H + vi * 8, gd[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < NVMAX; ++i) {
                const int vi = lane + 32 * i;
                if (vi < nv) {
                    float g[8], a[8];
                    load8(gamma + vi * 8, g);
                    if (drop_on_input && drop.thresh16) {
                        // y = dropout(LN(x)) (embedding modules): the incoming grad is masked first
                        uint32_t hw[4];
                rng_words8(seed, drop.stream, (size_t)row * H + vi * 8, hw);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t h = hw[j];
                            gd[i][2 * j] = ((h & 0xffffu) >= drop.thresh16) ? gd[i][2 * j] * drop.scale : 0.f;
                            gd[i][2 * j + 1] = ((h >> 16) >= drop.thresh16) ? gd[i][2 * j + 1] * drop.scale : 0.f;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        xh[i][j] = (xh[i][j] - mu) * rs;
                        a[j] = gd[i][j] * xh[i][j];
                    }
                    store8(sG + (size_t)warp * H + vi * 8, a);
                    store8(sB + (size_t)warp * H + vi * 8, gd[i]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        gd[i][j] *= g[j];
                        s1 += gd[i][j];
                        s2 += gd[i][j] * xh[i][j];
                    }
                }
            }
            s1 = warp_sum(s1) / (float)H;
            s2 = warp_sum(s2) / (float)H;
#pragma unroll
            for (int i = 0; i < NVMAX; ++i) {
                const int vi = lane + 32 * i;
                if (vi < nv) {
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = rs * (gd[i][j] - s1 - xh[i][j] * s2);
                    if (dx) store8(dx + (size_t)row * H + vi * 8, o);
                    if (dz) {
                        if (drop.thresh16 && !drop_on_input) {
                            uint32_t hw[4];
                rng_words8(seed, drop.stream, (size_t)row * H + vi * 8, hw);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint32_t h = hw[j];
                                o[2 * j] = ((h & 0xffffu) >= drop.thresh16) ? o[2 * j] * drop.scale : 0.f;
                                o[2 * j + 1] = ((h >> 16) >= drop.thresh16) ? o[2 * j + 1] * drop.scale : 0.f;
                            }
                        }
                        store8(dz + (size_t)row * H + vi * 8, o);
                    }
                    if (dbias) {
                        // column sums of the bf16-rounded dz: exactly what the wgrad GEMM consumes
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = __bfloat162float(__float2bfloat16(o[j]));
                        store8(sZ + (size_t)warp * H + vi * 8, o);
                    }
                }
            }
        }
        __syncthreads();
        const int nrows = min(rows_per_pass, r_end - base);
        if (c0 < H) {
            for (int r = 0; r < nrows; ++r) {
                const float2 a = *reinterpret_cast<const float2*>(sG + (size_t)r * H + c0);
                const float2 bb = *reinterpret_cast<const float2*>(sB + (size_t)r * H + c0);
                cg0 += a.x; cg1 += a.y; cb0 += bb.x; cb1 += bb.y;
                if (dbias) {
                    const float2 z = *reinterpret_cast<const float2*>(sZ + (size_t)r * H + c0);
                    cz0 += z.x; cz1 += z.y;
                }
            }
        }
        __syncthreads();  // staging is rewritten by the next pass
    }
    if (c0 < H && r_begin < r_end) {
        if (dgamma) { atomicAdd(dgamma + c0, cg0); atomicAdd(dgamma + c0 + 1, cg1); }
        if (dbeta) { atomicAdd(dbeta + c0, cb0); atomicAdd(dbeta + c0 + 1, cb1); }
        if (dbias) { atomicAdd(dbias + c0, cz0); atomicAdd(dbias + c0 + 1, cz1); }
    }
}

// ---------------------------------------------------------------------------------------
// out[n] += sum_m x[m, n]   (bias gradients of the GEMMs whose dY comes from another GEMM /
// the attention backward). Block = 8 warps x 256 columns; rows strided over blockIdx.y.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ x, int ldx, float* __restrict__ out, int M, int N) {
    pdl_sync();
    __shared__ float red[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 256 + lane * 8;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    if (col < N) {
        for (int row = blockIdx.y * 8 + warp; row < M; row += gridDim.y * 8) {
            float v[8];
            load8(x + (size_t)row * ldx + col, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] += v[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = a[j];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        atomicAdd(out + c, s);
    }
}

// fp32 -> bf16 cast (n multiple of 8), grid-stride.
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, size_t nvec) {
    pdl_sync();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec;
         i += (size_t)gridDim.x * blockDim.x) {
        float f[8];
        load8(x + i * 8, f);
        store8(y + i * 8, f);
    }
}

// dst[i] = bf16( f32(dst[i]) + sum_k f32(srcs[k][i]) ), sources added in the order given: the reduce step of the
// copy-engine all-reduce of a gradient bucket slice (train.py PeerBuckets). n multiple of 8.
struct SliceSrcs {
    const bf16* p[B200U_MAX_PEERS];
};
__global__ void slice_sum_bf16_kernel(bf16* __restrict__ dst, SliceSrcs srcs, int nsrc, size_t nvec) {
    pdl_sync();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec;
         i += (size_t)gridDim.x * blockDim.x) {
        float acc[8];
        load8(dst + i * 8, acc);
#pragma unroll 4
        for (int k = 0; k < nsrc; ++k) {
            float f[8];
            load8(srcs.p[k] + i * 8, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
        store8(dst + i * 8, acc);
    }
}

// out = dy * gelu_erf'(u)  (backward of the standalone dense+GELU transforms of the heads)
__global__ void dgelu_mul_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ u,
                                 bf16* __restrict__ out, size_t nvec) {
    pdl_sync();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec;
         i += (size_t)gridDim.x * blockDim.x) {
        float a[8], b[8];
        load8(dy + i * 8, a);
        load8(u + i * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] *= gelu_erf_grad(b[j]);
        store8(out + i * 8, a);
    }
}

// ---------------------------------------------------------------------------------------
// K1 gather: out[b, j, :] = (idx < T ? txt[b, idx] : img[b, idx - T]),  idx = gather_index[b, j]
// Pure 16-byte row copies -> bit-exact with torch.gather on the concatenation, padded
// positions included. Index is read once per row, never expanded.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_rows_kernel(const bf16* __restrict__ txt, const bf16* __restrict__ img,
                   const long long* __restrict__ gidx, bf16* __restrict__ out, int B, int T, int R,
                   int L, int H, unsigned* __restrict__ err) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B * L) return;
    const int b = row / L;
    long long idx = gidx[row];
    if (idx < 0 || idx >= T + R) {   // torch.gather raises: flag it and read row 0 instead of out of bounds
        if (lane == 0 && err) atomicOr(err, (unsigned)ERR_GATHER_INDEX);
        idx = 0;
    }
    const bf16* src = (idx < T) ? txt + ((size_t)b * T + idx) * H : img + ((size_t)b * R + (idx - T)) * H;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(out + (size_t)row * H);
    for (int v = lane; v < (H >> 3); v += 32) d4[v] = s4[v];
}

// Backward of the gather = scatter-add; done as a deterministic inverse scan: the warp that owns
// source row s of sample b sums every dout[b, j] with gather_index[b, j] == s (fp32), so duplicate
// indices (the identity tail of get_gather_index) need no atomics.
__global__ void __launch_bounds__(256)
gather_rows_bwd_kernel(const bf16* __restrict__ dout, const long long* __restrict__ gidx,
                       bf16* __restrict__ dtxt, bf16* __restrict__ dimg, int B, int T, int R, int L,
                       int H) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int srow = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int S = T + R;
    if (srow >= B * S) return;
    const int b = srow / S, s = srow - b * S;
    const int nv = H >> 3;
    float acc[LN_MAXV][8];
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int j0 = 0; j0 < L; j0 += 32) {
        const int j = j0 + lane;
        const bool hit = (j < L) && (gidx[(size_t)b * L + j] == (long long)s);
        unsigned m = __ballot_sync(0xffffffffu, hit);
        while (m) {
            const int jj = j0 + __ffs(m) - 1;
            m &= m - 1;
#pragma unroll
            for (int i = 0; i < LN_MAXV; ++i) {
                const int vi = lane + 32 * i;
                if (vi < nv) {
                    float d[8];
                    load8(dout + ((size_t)b * L + jj) * H + vi * 8, d);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[i][k] += d[k];
                }
            }
        }
    }
    bf16* dst = (s < T) ? dtxt + ((size_t)b * T + s) * H : dimg + ((size_t)b * R + (s - T)) * H;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nv) store8(dst + vi * 8, acc[i]);
    }
}

static DropoutCfg make_drop(const b200u_dropout_t& d) {
    DropoutCfg c;
    c.seed_ptr = d.seed_ptr;
    c.stream = d.stream;
    c.thresh16 = (uint32_t)(d.p * 65536.0f + 0.5f);
    c.scale = 1.0f / (1.0f - d.p);
    return c;
}

}  // namespace b200u

using namespace b200u;

#define CHECK_H(H) \
    B200U_CHECK_ARG((H) > 0 && (H) % 8 == 0 && (H) <= 8 * 32 * LN_MAXV, "hidden size %d unsupported (need H%%8==0, H<=1024)", (H))

extern "C" int b200u_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta,
                                   void* y, int y_dtype, float* mean, float* rstd, int M, int H,
                                   float eps, const b200u_dropout_t* drop, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CHECK_H(H);
    B200U_CHECK_ARG(M >= 0 && x && gamma && beta && y, "layernorm_fwd: null pointer");
    if (M == 0) return B200U_OK;
    b200u_dropout_t nodrop = {nullptr, 0, 0.f};
    DropoutCfg dc = make_drop(drop ? *drop : nodrop);
    B200U_CHECK_ARG(dc.thresh16 == 0 || dc.seed_ptr, "layernorm_fwd: dropout needs seed_ptr");
    const int grid = (M + 7) / 8;
    if (x_dtype == B200U_BF16 && y_dtype == B200U_BF16)
        launch_k(layernorm_fwd_kernel<bf16, bf16>, dim3(grid), dim3(256), 0, stream, (const bf16*)x, gamma, beta, (bf16*)y, mean, rstd, M, H, eps, dc);
    else if (x_dtype == B200U_F32 && y_dtype == B200U_BF16)
        launch_k(layernorm_fwd_kernel<float, bf16>, dim3(grid), dim3(256), 0, stream, (const float*)x, gamma, beta, (bf16*)y, mean, rstd, M, H, eps, dc);
    else if (x_dtype == B200U_F32 && y_dtype == B200U_F32)
        launch_k(layernorm_fwd_kernel<float, float>, dim3(grid), dim3(256), 0, stream, (const float*)x, gamma, beta, (float*)y, mean, rstd, M, H, eps, dc);
    else
        B200U_CHECK_ARG(false, "layernorm_fwd: unsupported dtype combination %d -> %d", x_dtype, y_dtype);
    B200U_CHECK_LAUNCH("layernorm_fwd");
    return B200U_OK;
}

extern "C" int b200u_layernorm_bwd(const void* dy, const void* x, int x_dtype, const float* mean,
                                   const float* rstd, const float* gamma, void* dx, void* dz,
                                   float* dgamma, float* dbeta, float* dbias, int M, int H,
                                   const b200u_dropout_t* drop, int drop_on_input,
                                   b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CHECK_H(H);
    B200U_CHECK_ARG(dy && x && mean && rstd && gamma, "layernorm_bwd: null pointer");
    if (M == 0) return B200U_OK;
    b200u_dropout_t nodrop = {nullptr, 0, 0.f};
    DropoutCfg dc = make_drop(drop ? *drop : nodrop);
    B200U_CHECK_ARG(dc.thresh16 == 0 || dc.seed_ptr, "layernorm_bwd: dropout needs seed_ptr");
    // one row per warp per pass; as many rows per pass as the staging tile [3][rows][H] fp32 allows
    int rows_per_pass = (int)((size_t)200 * 1024 / ((size_t)3 * H * sizeof(float)));
    if (rows_per_pass > LNB_WARPS) rows_per_pass = LNB_WARPS;
    int grid = (M + rows_per_pass - 1) / rows_per_pass;
    if (grid > num_sms()) grid = num_sms();
    const int rows_per_cta = (M + grid - 1) / grid;
    grid = (M + rows_per_cta - 1) / rows_per_cta;
    const size_t smem = (size_t)3 * rows_per_pass * H * sizeof(float);
    if (int rc = ensure_dyn_smem((const void*)layernorm_bwd_kernel<bf16, 3>, 200 * 1024)) return rc;
    if (int rc = ensure_dyn_smem((const void*)layernorm_bwd_kernel<bf16, 4>, 200 * 1024)) return rc;
    if (int rc = ensure_dyn_smem((const void*)layernorm_bwd_kernel<float, 3>, 200 * 1024)) return rc;
    if (int rc = ensure_dyn_smem((const void*)layernorm_bwd_kernel<float, 4>, 200 * 1024)) return rc;
    const bool narrow = H <= 768;
#define LNB_LAUNCH(TX, NV) \
    launch_k(layernorm_bwd_kernel<TX, NV>, dim3(grid), dim3(LNB_THREADS), smem, stream, (const bf16*)dy, (const TX*)x, mean, rstd, gamma, (bf16*)dx, (bf16*)dz, dgamma, dbeta, dbias, M, H, rows_per_cta, rows_per_pass, dc, drop_on_input)
    if (x_dtype == B200U_BF16) {
        if (narrow) LNB_LAUNCH(bf16, 3); else LNB_LAUNCH(bf16, 4);
    } else if (x_dtype == B200U_F32) {
        if (narrow) LNB_LAUNCH(float, 3); else LNB_LAUNCH(float, 4);
    }
#undef LNB_LAUNCH
    else
        B200U_CHECK_ARG(false, "layernorm_bwd: unsupported x dtype %d", x_dtype);
    B200U_CHECK_LAUNCH("layernorm_bwd");
    return B200U_OK;
}

extern "C" int b200u_colsum_accum(const void* x, int ldx, float* out, int M, int N,
                                  b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(x && out && N % 8 == 0 && ldx % 8 == 0, "colsum_accum: bad arguments");
    if (M == 0) return B200U_OK;
    const int gx = (N + 255) / 256;
    int gy = (2 * num_sms() + gx - 1) / gx;
    if (gy > (M + 7) / 8) gy = (M + 7) / 8;
    if (gy < 1) gy = 1;
    launch_k(colsum_kernel, dim3(dim3(gx, gy)), dim3(256), 0, stream, (const bf16*)x, ldx, out, M, N);
    B200U_CHECK_LAUNCH("colsum_accum");
    return B200U_OK;
}

extern "C" int b200u_cast_f32_to_bf16(const float* x, void* y, size_t n, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(x && y && n % 8 == 0, "cast_f32_to_bf16: n must be a multiple of 8");
    if (n == 0) return B200U_OK;
    const size_t nvec = n / 8;
    size_t grid = (nvec + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    if (grid > cap) grid = cap;
    launch_k(cast_f32_bf16_kernel, dim3((int)grid), dim3(256), 0, stream, x, (bf16*)y, nvec);
    B200U_CHECK_LAUNCH("cast_f32_to_bf16");
    return B200U_OK;
}

extern "C" int b200u_slice_sum_bf16(void* dst, const void* const* srcs, int nsrc, size_t n, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(dst && (srcs || nsrc == 0) && nsrc >= 0 && nsrc <= B200U_MAX_PEERS && n % 8 == 0 &&
                        ((uintptr_t)dst & 15) == 0,
                    "slice_sum_bf16: need <= %d 16-byte aligned sources and n %% 8 == 0", B200U_MAX_PEERS);
    if (n == 0 || nsrc == 0) return B200U_OK;
    SliceSrcs ss = {};
    for (int k = 0; k < nsrc; ++k) {
        B200U_CHECK_ARG(srcs[k] && ((uintptr_t)srcs[k] & 15) == 0, "slice_sum_bf16: source %d null or misaligned", k);
        ss.p[k] = (const bf16*)srcs[k];
    }
    const size_t nvec = n / 8;
    size_t grid = (nvec + 255) / 256;
    const size_t cap = (size_t)num_sms() * 2;   // a side-stream kernel beside the backward GEMMs: keep it small
    if (grid > cap) grid = cap;
    launch_k(slice_sum_bf16_kernel, dim3((int)grid), dim3(256), 0, stream, (bf16*)dst, ss, nsrc, nvec);
    B200U_CHECK_LAUNCH("slice_sum_bf16");
    return B200U_OK;
}

extern "C" int b200u_dgelu_mul(const void* dy, const void* u, void* out, size_t n, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(dy && u && out && n % 8 == 0, "dgelu_mul: n must be a multiple of 8");
    if (n == 0) return B200U_OK;
    const size_t nvec = n / 8;
    size_t grid = (nvec + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    if (grid > cap) grid = cap;
    launch_k(dgelu_mul_kernel, dim3((int)grid), dim3(256), 0, stream, (const bf16*)dy, (const bf16*)u, (bf16*)out, nvec);
    B200U_CHECK_LAUNCH("dgelu_mul");
    return B200U_OK;
}

extern "C" int b200u_gather_rows(const void* txt, const void* img, const long long* gather_index,
                                 void* out, int B, int T, int R, int L, int H, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(txt && img && gather_index && out && H % 8 == 0, "gather_rows: bad arguments");
    if (B * L == 0) return B200U_OK;
    launch_k(gather_rows_kernel, dim3((B * L + 7) / 8), dim3(256), 0, stream, (const bf16*)txt, (const bf16*)img, gather_index, (bf16*)out, B, T, R, L, H, dev_err_ptr());
    B200U_CHECK_LAUNCH("gather_rows");
    return B200U_OK;
}

extern "C" int b200u_gather_rows_bwd(const void* dout, const long long* gather_index, void* dtxt,
                                     void* dimg, int B, int T, int R, int L, int H,
                                     b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CHECK_H(H);
    B200U_CHECK_ARG(dout && gather_index && dtxt && dimg, "gather_rows_bwd: null pointer");
    if (B * (T + R) == 0) return B200U_OK;
    launch_k(gather_rows_bwd_kernel, dim3((B * (T + R) + 7) / 8), dim3(256), 0, stream, (const bf16*)dout, gather_index, (bf16*)dtxt, (bf16*)dimg, B, T, R, L, H);
    B200U_CHECK_LAUNCH("gather_rows_bwd");
    return B200U_OK;
}
