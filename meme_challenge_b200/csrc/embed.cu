// K0 / K2 — UNITER text and image embedding kernels (sm_100a).
//
//  txt_embed_fwd : LN(word[ids] + pos[position_ids] + type[type_ids]) -> dropout
//                  (model/model.py:232-245), one warp per token, tables read in fp32.
//  img_embed_fwd : LN( LN_img(a) + LN_pos(pos7·W_posᵀ + b_pos) + type[type_ids] ) -> dropout
//                  (model/model.py:261-272; `a` = img_linear output from the tcgen05 GEMM). The
//                  K=7 pos_linear is 7 FMAs per element on CUDA cores, fused here with the
//                  three LayerNorms so the 768-wide row is read once and written once.
//  embedding_scatter_add / pos_linear_wgrad : the parameter-gradient halves of their backward
//                  (the LayerNorm halves reuse layernorm_bwd from rowops.cu).
#include "../../include/b200u.h"
#include "common.cuh"

#include <mutex>

namespace b200u {

constexpr int EMB_MAXV = 4;  // H <= 1024

__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8f(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void st8b(bf16* p, const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
    o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = o;
}

__device__ __forceinline__ void stats(const float (&x)[EMB_MAXV][8], int nv, int lane, int H,
                                      float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < EMB_MAXV; ++i)
        if (lane + 32 * i < nv)
#pragma unroll
            for (int j = 0; j < 8; ++j) s += x[i][j];
    mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < EMB_MAXV; ++i)
        if (lane + 32 * i < nv)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float d = x[i][j] - mean;
                q += d * d;
            }
    rstd = rsqrtf(warp_sum(q) / (float)H + eps);
}

__device__ __forceinline__ void apply_dropout8(float (&o)[8], uint64_t seed, const DropoutCfg& drop,
                                               size_t elem0) {
    if (!drop.thresh16) return;
    uint32_t hw[4];
    rng_words8(seed, drop.stream, elem0, hw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t h = hw[j];
        o[2 * j] = ((h & 0xffffu) >= drop.thresh16) ? o[2 * j] * drop.scale : 0.f;
        o[2 * j + 1] = ((h >> 16) >= drop.thresh16) ? o[2 * j + 1] * drop.scale : 0.f;
    }
}

__global__ void __launch_bounds__(256)
txt_embed_fwd_kernel(const long long* __restrict__ ids, const long long* __restrict__ pos_ids,
                     int pos_batch_stride, const long long* __restrict__ type_ids,
                     const float* __restrict__ word, const float* __restrict__ pos,
                     const float* __restrict__ type, const float* __restrict__ gamma,
                     const float* __restrict__ beta, bf16* __restrict__ out, float* __restrict__ sum_out,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int n, int T, int H,
                     int vocab_rows, int pos_rows, int type_rows, unsigned* __restrict__ err,
                     float eps, DropoutCfg drop) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tok >= n) return;
    const int b = tok / T, t = tok - b * T;
    long long wid = ids[tok];
    long long pid = pos_ids[(size_t)b * pos_batch_stride + t];
    long long tid = type_ids ? type_ids[tok] : 0;
    // out-of-range ids (nn.Embedding raises): flag them and read row 0 instead of out of bounds
    unsigned bad = 0;
    if (wid < 0 || wid >= vocab_rows) { bad |= ERR_WORD_ID; wid = 0; }
    if (pid < 0 || pid >= pos_rows) { bad |= ERR_POS_ID; pid = 0; }
    if (tid < 0 || tid >= type_rows) { bad |= ERR_TYPE_ID; tid = 0; }
    if (bad && lane == 0 && err) atomicOr(err, bad);
    const int nv = H >> 3;
    float x[EMB_MAXV][8];
#pragma unroll
    for (int i = 0; i < EMB_MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nv) {
            float w[8], p[8], ty[8];
            ld8f(word + (size_t)wid * H + vi * 8, w);
            ld8f(pos + (size_t)pid * H + vi * 8, p);
            ld8f(type + (size_t)tid * H + vi * 8, ty);
#pragma unroll
            for (int j = 0; j < 8; ++j) x[i][j] = (w[j] + p[j]) + ty[j];  // model.py:240-242 order
            if (sum_out) st8f(sum_out + (size_t)tok * H + vi * 8, x[i]);
        }
    }
    float mean, rstd;
    stats(x, nv, lane, H, eps, mean, rstd);
    if (lane == 0) {
        if (mean_out) mean_out[tok] = mean;
        if (rstd_out) rstd_out[tok] = rstd;
    }
    const uint64_t seed = load_seed(drop);
#pragma unroll
    for (int i = 0; i < EMB_MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nv) {
            float g[8], be[8], o[8];
            ld8f(gamma + vi * 8, g);
            ld8f(beta + vi * 8, be);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (x[i][j] - mean) * rstd * g[j] + be[j];
            apply_dropout8(o, seed, drop, (size_t)tok * H + vi * 8);
            st8b(out + (size_t)tok * H + vi * 8, o);
        }
    }
}

// stats layout: [6, n] = (mean_img, rstd_img, mean_pos, rstd_pos, mean, rstd) rows
__global__ void __launch_bounds__(256)
img_embed_fwd_kernel(const float* __restrict__ a, const float* __restrict__ pos7,
                     const float* __restrict__ Wpos, const float* __restrict__ bpos,
                     const long long* __restrict__ type_ids, const float* __restrict__ type,
                     const float* __restrict__ g_img, const float* __restrict__ b_img,
                     const float* __restrict__ g_pos, const float* __restrict__ b_pos,
                     const float* __restrict__ g, const float* __restrict__ be,
                     bf16* __restrict__ out, float* __restrict__ p_out, float* __restrict__ s_out,
                     float* __restrict__ stats_out, int n, int H, int type_rows, unsigned* __restrict__ err,
                     float eps, DropoutCfg drop) {
    pdl_sync();
    // Everything that does not depend on the row is staged once per CTA with 16-byte loads in ONE round trip:
    // pos_linear's weight [H][7] (a lane's 8 consecutive features are 56 contiguous floats) and bias, the three
    // LayerNorms' gamma / beta and token-type rows 0 and 1. The row's own loads are issued before the staging so
    // both latencies overlap; after the barrier the kernel only touches shared memory until its stores.
    extern __shared__ __align__(16) float sWp[];   // [H * 7] weight | [H] bias | 6 x [H] LN vectors | 2 x [H] type rows
    float* sVec = sWp + (size_t)H * 8;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const bool valid = row < n;
    const int nv = H >> 3;
    float pf[7];
    long long tid = 1;  // model.py:313-314 default ones
    float xa[EMB_MAXV][8], xp[EMB_MAXV][8];
    if (valid) {
#pragma unroll
        for (int c = 0; c < 7; ++c) pf[c] = pos7[(size_t)row * 7 + c];
        if (type_ids) tid = type_ids[row];
        if (tid < 0 || tid >= type_rows) {
            if (lane == 0 && err) atomicOr(err, (unsigned)ERR_TYPE_ID);
            tid = 0;
        }
#pragma unroll
        for (int i = 0; i < EMB_MAXV; ++i)
            if (lane + 32 * i < nv) ld8f(a + (size_t)row * H + (lane + 32 * i) * 8, xa[i]);
    }
    {
        const int nvec = (H * 7) >> 2, hv = H >> 2;   // H % 8 == 0
        for (int i = threadIdx.x; i < nvec; i += blockDim.x)
            reinterpret_cast<float4*>(sWp)[i] = reinterpret_cast<const float4*>(Wpos)[i];
        const float* vecs[8] = {g_img, b_img, g_pos, b_pos, g, be, type, type + H};
        for (int i = threadIdx.x; i < hv; i += blockDim.x) {
            reinterpret_cast<float4*>(sWp + H * 7)[i] = reinterpret_cast<const float4*>(bpos)[i];
#pragma unroll
            for (int v = 0; v < 8; ++v)
                reinterpret_cast<float4*>(sVec + (size_t)v * H)[i] = reinterpret_cast<const float4*>(vecs[v])[i];
        }
    }
    __syncthreads();
    if (!valid) return;
    const float* ty_row = (tid >= 0 && tid < 2) ? sVec + (size_t)(6 + tid) * H : type + (size_t)tid * H;

#pragma unroll
    for (int i = 0; i < EMB_MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nv) {
            float wflat[56], bv[8];
#pragma unroll
            for (int q = 0; q < 14; ++q) {
                const float4 t4 = *reinterpret_cast<const float4*>(sWp + (size_t)vi * 56 + 4 * q);
                wflat[4 * q] = t4.x; wflat[4 * q + 1] = t4.y; wflat[4 * q + 2] = t4.z; wflat[4 * q + 3] = t4.w;
            }
            ld8f(sWp + H * 7 + vi * 8, bv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 7; ++c) acc = fmaf(pf[c], wflat[j * 7 + c], acc);
                xp[i][j] = acc + bv[j];
            }
            if (p_out) st8f(p_out + (size_t)row * H + vi * 8, xp[i]);
        }
    }
    float m1, r1, m2, r2, m3, r3;
    stats(xa, nv, lane, H, eps, m1, r1);
    stats(xp, nv, lane, H, eps, m2, r2);
#pragma unroll
    for (int i = 0; i < EMB_MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nv) {
            float gi[8], bi[8], gp[8], bp[8], ty[8];
            ld8f(sVec + vi * 8, gi); ld8f(sVec + H + vi * 8, bi);
            ld8f(sVec + 2 * H + vi * 8, gp); ld8f(sVec + 3 * H + vi * 8, bp);
            ld8f(ty_row + vi * 8, ty);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t1 = (xa[i][j] - m1) * r1 * gi[j] + bi[j];
                const float t2 = (xp[i][j] - m2) * r2 * gp[j] + bp[j];
                xa[i][j] = (t1 + t2) + ty[j];  // model.py:269 order
            }
            if (s_out) st8f(s_out + (size_t)row * H + vi * 8, xa[i]);
        }
    }
    stats(xa, nv, lane, H, eps, m3, r3);
    if (lane == 0 && stats_out) {
        float* st = stats_out + row;
        st[0] = m1; st[(size_t)n] = r1; st[(size_t)2 * n] = m2; st[(size_t)3 * n] = r2;
        st[(size_t)4 * n] = m3; st[(size_t)5 * n] = r3;
    }
    const uint64_t seed = load_seed(drop);
#pragma unroll
    for (int i = 0; i < EMB_MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nv) {
            float gg[8], bb[8], o[8];
            ld8f(sVec + 4 * H + vi * 8, gg);
            ld8f(sVec + 5 * H + vi * 8, bb);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (xa[i][j] - m3) * r3 * gg[j] + bb[j];
            apply_dropout8(o, seed, drop, (size_t)row * H + vi * 8);
            st8b(out + (size_t)row * H + vi * 8, o);
        }
    }
}

// table_grad[ids[r], :] += d[r, :]   (nn.Embedding backward; rows with ids == padding_idx skipped)
// One warp per (position t, 256-column chunk, segment of 8 samples): it loads the segment's rows b*T + t together
// and keeps the running sum of consecutive rows with the SAME id in registers, flushing (two 16-byte vector
// reductions per lane) only when the id changes. Position ids repeat down the batch (one flush per segment instead
// of 8 colliding atomics per element) and so do [CLS] / [SEP]; distinct ids cost one flush each.
__global__ void __launch_bounds__(64)
embedding_scatter_add_kernel(const bf16* __restrict__ d, const long long* __restrict__ ids,
                             int ids_batch_stride, int T, long long const_id, float* __restrict__ table_grad,
                             int n, int H, long long padding_idx, int nchunk, int nseg, long long rows,
                             unsigned* __restrict__ err) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int c = w % nchunk;
    w /= nchunk;
    const int seg = w % nseg, t = w / nseg;
    if (t >= T) return;
    const int B = n / T;
    const int b0 = seg * 8;
    const int col = c * 256 + lane * 8;
    if (col >= H) return;
    const int bb = b0 + (lane & 7);
    long long my_id = const_id;
    if (ids && bb < B) my_id = ids[(size_t)bb * ids_batch_stride + t];
    if (my_id != padding_idx && (my_id < 0 || my_id >= rows)) {   // never add outside the table: flag and skip the row
        if (err && c == 0) atomicOr(err, (unsigned)ERR_SCATTER_ID);
        my_id = padding_idx;
    }
    uint4 u[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
        u[r] = (b0 + r < B) ? *reinterpret_cast<const uint4*>(d + ((size_t)(b0 + r) * T + t) * H + col)
                            : make_uint4(0, 0, 0, 0);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    long long cur = padding_idx;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const long long id = __shfl_sync(0xffffffffu, my_id, r);
        if (b0 + r < B) {
            if (id != cur) {
                if (cur != padding_idx) {
                    float* dst = table_grad + (size_t)cur * H + col;
                    red_add_v4(dst, acc[0], acc[1], acc[2], acc[3]);
                    red_add_v4(dst + 4, acc[4], acc[5], acc[6], acc[7]);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = 0.f;
                cur = id;
            }
            const uint32_t* up = &u[r].x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = unpack_bf16(up[k]);
                acc[2 * k] += f.x;
                acc[2 * k + 1] += f.y;
            }
        }
    }
    if (cur != padding_idx) {
        float* dst = table_grad + (size_t)cur * H + col;
        red_add_v4(dst, acc[0], acc[1], acc[2], acc[3]);
        red_add_v4(dst + 4, acc[4], acc[5], acc[6], acc[7]);
    }
}

// Deterministic variant for rows sorted by id: ids_sorted[p] ascending, perm[p] = source row of the
// p-th sorted entry (stable sort). One warp per RUN of equal ids: it adds the run's rows in sorted order
// into registers and does one plain read-modify-write of table_grad[id, :] -- no atomics, so identical
// inputs give bit-identical tables on every rank (data-parallel sparse exchange of the word-embedding
// gradient, train.py).
// One warp per (run, 256-column chunk): lane l owns 8 consecutive columns. Frequent tokens ([CLS], [SEP] occur in
// every sample) form runs of B * world rows, so the row loop keeps 8 row loads in flight (their source rows come
// from one coalesced read of `perm`) and still adds them in sorted order: the sum order is fixed.
__global__ void __launch_bounds__(256)
embedding_segment_add_kernel(const bf16* __restrict__ d, const long long* __restrict__ ids_sorted,
                             const long long* __restrict__ perm, float* __restrict__ table_grad, int n,
                             int H, long long padding_idx, int nchunk, long long rows,
                             unsigned* __restrict__ err, double* __restrict__ sumsq, int sumsq_slots) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int p = w / nchunk, c = w - p * nchunk;
    if (p >= n) return;
    const long long id = ids_sorted[p];
    if (id == padding_idx) return;
    if (id < 0 || id >= rows) {
        if (err && lane == 0 && c == 0) atomicOr(err, (unsigned)ERR_SCATTER_ID);
        return;
    }
    if (p > 0 && ids_sorted[p - 1] == id) return;  // not the first entry of its run
    const int col = c * 256 + lane * 8;
    const bool active = col < H;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    int q = p;
    while (q < n) {
        // lanes 0..7 probe the next 8 sorted entries; the warp learns how many still belong to this run
        const int qq = q + (lane & 7);
        const bool mine = qq < n && ids_sorted[qq] == id;
        const long long src = mine ? perm[qq] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, mine) & 0xffu;
        const int cnt = __ffs(~m & 0x1ffu) - 1;   // set bits are contiguous from bit 0 (sorted ids): run length, 0..8
        if (cnt == 0) break;
        uint4 u[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const long long sr = __shfl_sync(0xffffffffu, src, r);
            u[r] = (r < cnt && active) ? *reinterpret_cast<const uint4*>(d + (size_t)sr * H + col) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r < cnt) {
                const uint32_t* up = &u[r].x;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = unpack_bf16(up[k]);
                    acc[2 * k] += f.x;
                    acc[2 * k + 1] += f.y;
                }
            }
        }
        if (cnt < 8) break;
        q += 8;
    }
    float sq = 0.f;
    if (active) {
        float4* o = reinterpret_cast<float4*>(table_grad + (size_t)id * H + col);
        float4 a = o[0], b = o[1];
        a.x += acc[0]; a.y += acc[1]; a.z += acc[2]; a.w += acc[3];
        b.x += acc[4]; b.y += acc[5]; b.z += acc[6]; b.w += acc[7];
        o[0] = a; o[1] = b;
        sq = (a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w) + (b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w);
    }
    if (sumsq) {   // the table's share of the global gradient norm, valid when the table was zero before this launch
        sq = warp_sum(sq);
        // thousands of warps: spread over the caller's slots, same-address atomics would serialise in L2
        if (lane == 0) atomicAdd(sumsq + (blockIdx.x % (unsigned)sumsq_slots), (double)sq);
    }
}

// dW[h, c] += sum_r dp[r, h] * pos7[r, c]   (pos_linear weight grad, K = 7).
// CTA = 64 features x 128 rows: thread (h, row quarter) sums its 32 rows with 8 loads in flight, the four quarters meet
// in shared memory and one thread per feature issues the 7 atomics.
__global__ void __launch_bounds__(256)
pos_linear_wgrad_kernel(const bf16* __restrict__ dp, const float* __restrict__ pos7,
                        float* __restrict__ dW, int n, int H) {
    pdl_sync();
    const int hl = threadIdx.x & 63, sub = threadIdx.x >> 6;
    const int h = blockIdx.x * 64 + hl;
    const int r0 = blockIdx.y * 128;
    const int cnt = min(128, n - r0);
    __shared__ float sp[128][8];
    __shared__ float red[3][64][7];
    for (int i = threadIdx.x; i < 128 * 7; i += blockDim.x) {
        const int rr = i / 7, c = i - rr * 7;
        sp[rr][c] = rr < cnt ? pos7[(size_t)(r0 + rr) * 7 + c] : 0.f;
    }
    __syncthreads();
    float acc[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) acc[c] = 0.f;
    if (h < H) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = sub * 32 + g * 8 + u;
                v[u] = rr < cnt ? __bfloat162float(dp[(size_t)(r0 + rr) * H + h]) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = sub * 32 + g * 8 + u;
#pragma unroll
                for (int c = 0; c < 7; ++c) acc[c] = fmaf(v[u], sp[rr][c], acc[c]);
            }
        }
    }
    if (sub > 0)
#pragma unroll
        for (int c = 0; c < 7; ++c) red[sub - 1][hl][c] = acc[c];
    __syncthreads();
    if (sub == 0 && h < H)
#pragma unroll
        for (int c = 0; c < 7; ++c)
            atomicAdd(dW + (size_t)h * 7 + c, acc[c] + red[0][hl][c] + red[1][hl][c] + red[2][hl][c]);
}

static DropoutCfg make_drop(const b200u_dropout_t* d) {
    DropoutCfg c;
    c.seed_ptr = d ? d->seed_ptr : nullptr;
    c.stream = d ? d->stream : 0;
    const float p = d ? d->p : 0.f;
    c.thresh16 = (uint32_t)(p * 65536.0f + 0.5f);
    c.scale = 1.0f / (1.0f - p);
    return c;
}

}  // namespace b200u

using namespace b200u;

#define CHECK_H(H) \
    B200U_CHECK_ARG((H) > 0 && (H) % 8 == 0 && (H) <= 8 * 32 * EMB_MAXV, "hidden size %d unsupported (need H%%8==0, H<=1024)", (H))

extern "C" int b200u_txt_embed_fwd(const long long* input_ids, const long long* position_ids,
                                   int pos_batch_stride, const long long* type_ids, const float* word,
                                   const float* pos, const float* type, const float* gamma,
                                   const float* beta, void* out, float* sum_out, float* mean,
                                   float* rstd, int B, int T, int H, int vocab_rows, int pos_rows,
                                   int type_rows, float eps, const b200u_dropout_t* drop,
                                   b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CHECK_H(H);
    B200U_CHECK_ARG(input_ids && position_ids && word && pos && type && gamma && beta && out, "txt_embed_fwd: null pointer");
    const int n = B * T;
    if (n == 0) return B200U_OK;
    DropoutCfg dc = make_drop(drop);
    B200U_CHECK_ARG(dc.thresh16 == 0 || dc.seed_ptr, "txt_embed_fwd: dropout needs seed_ptr");
    launch_k(txt_embed_fwd_kernel, dim3((n + 7) / 8), dim3(256), 0, stream, input_ids, position_ids, pos_batch_stride, type_ids, word, pos, type, gamma, beta, (bf16*)out, sum_out, mean, rstd, n, T, H, vocab_rows, pos_rows, type_rows, dev_err_ptr(), eps, dc);
    B200U_CHECK_LAUNCH("txt_embed_fwd");
    return B200U_OK;
}

extern "C" int b200u_img_embed_fwd(const float* a, const float* pos7, const float* Wpos,
                                   const float* bpos, const long long* type_ids, const float* type,
                                   const float* g_img, const float* b_img, const float* g_pos,
                                   const float* b_pos, const float* g, const float* b, void* out,
                                   float* p_out, float* s_out, float* stats_out, int n, int H,
                                   int type_rows, float eps, const b200u_dropout_t* drop,
                                   b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CHECK_H(H);
    B200U_CHECK_ARG(a && pos7 && Wpos && bpos && type && g_img && b_img && g_pos && b_pos && g && b && out, "img_embed_fwd: null pointer");
    B200U_CHECK_ARG(type_rows >= 2, "img_embed_fwd: the token-type table needs >= 2 rows (image regions default to type 1)");
    if (n == 0) return B200U_OK;
    DropoutCfg dc = make_drop(drop);
    B200U_CHECK_ARG(dc.thresh16 == 0 || dc.seed_ptr, "img_embed_fwd: dropout needs seed_ptr");
    const size_t smem = (size_t)16 * H * sizeof(float);
    {
        static std::mutex mu;
        static size_t set_for[64] = {};
        int dev = 0;
        B200U_CHECK_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 0 && dev < 64 && smem > set_for[dev]) {
            B200U_CHECK_CUDA(cudaFuncSetAttribute(img_embed_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set_for[dev] = smem;
        }
    }
    launch_k(img_embed_fwd_kernel, dim3((n + 7) / 8), dim3(256), smem, stream, a, pos7, Wpos, bpos, type_ids, type, g_img, b_img, g_pos, b_pos, g, b, (bf16*)out, p_out, s_out, stats_out, n, H, type_rows, dev_err_ptr(), eps, dc);
    B200U_CHECK_LAUNCH("img_embed_fwd");
    return B200U_OK;
}

extern "C" int b200u_embedding_scatter_add(const void* d, const long long* ids, int ids_batch_stride,
                                           int T, long long const_id, float* table_grad, int n, int H,
                                           long long padding_idx, long long rows, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(d && table_grad && H % 8 == 0 && T > 0 && n % T == 0, "embedding_scatter_add: bad arguments (n must be B * T)");
    if (n == 0) return B200U_OK;
    B200U_CHECK_ARG(((uintptr_t)table_grad & 15) == 0, "embedding_scatter_add: table_grad must be 16-byte aligned");
    const int nchunk = (H + 255) / 256, nseg = (n / T + 7) / 8;
    const long long warps = (long long)T * nseg * nchunk;
    launch_k(embedding_scatter_add_kernel, dim3((unsigned)((warps + 1) / 2)), dim3(64), 0, stream, (const bf16*)d, ids, ids_batch_stride, T, const_id, table_grad, n, H, padding_idx, nchunk, nseg, rows, dev_err_ptr());
    B200U_CHECK_LAUNCH("embedding_scatter_add");
    return B200U_OK;
}

extern "C" int b200u_embedding_segment_add(const void* d, const long long* ids_sorted, const long long* perm,
                                           float* table_grad, int n, int H, long long padding_idx,
                                           long long rows, double* sumsq, int sumsq_slots,
                                           b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(d && ids_sorted && perm && table_grad, "embedding_segment_add: null pointer");
    B200U_CHECK_ARG(H % 8 == 0, "embedding_segment_add: H must be a multiple of 8");
    B200U_CHECK_ARG(!sumsq || sumsq_slots >= 1, "embedding_segment_add: sumsq needs at least one slot");
    if (n == 0) return B200U_OK;
    const int nchunk = (H + 255) / 256;
    const long long warps = (long long)n * nchunk;
    launch_k(embedding_segment_add_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, stream, (const bf16*)d,
             ids_sorted, perm, table_grad, n, H, padding_idx, nchunk, rows, dev_err_ptr(), sumsq, sumsq_slots);
    B200U_CHECK_LAUNCH("embedding_segment_add_kernel");
    return B200U_OK;
}

extern "C" int b200u_pos_linear_wgrad(const void* dp, const float* pos7, float* dW, int n, int H,
                                      b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(dp && pos7 && dW, "pos_linear_wgrad: null pointer");
    if (n == 0) return B200U_OK;
    dim3 grid((H + 63) / 64, (n + 127) / 128);
    launch_k(pos_linear_wgrad_kernel, dim3(grid), dim3(256), 0, stream, (const bf16*)dp, pos7, dW, n, H);
    B200U_CHECK_LAUNCH("pos_linear_wgrad");
    return B200U_OK;
}
