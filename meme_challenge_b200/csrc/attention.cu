// K4 — fused masked-softmax self-attention over the joint text+region sequence (sm_100a).
//
// Replaces model/layer.py:80-100 (BertSelfAttention.forward after the Q/K/V projections):
//   scores = Q·Kᵀ ; scores = scores / sqrt(64) ; scores += additive_mask ; P = softmax(scores)
//   P = dropout(P) ; ctx = P·V ; heads merged back to [B, L, H]
// and its autograd backward. L <= 256 and head_dim == 64, so the whole K and V of one
// (sample, head) sit in shared memory and the [L, L] score matrix never touches HBM: a warp
// owns 16 query rows, keeps its score strip in registers (softmax by quad shuffles) and feeds
// bf16 probabilities straight back into the P·V tensor-core MMAs. Only the row log-sum-exp is
// saved; backward recomputes P and regenerates the dropout mask from the counter RNG.
//
// The joint sequence is at most 164 long with 64-wide heads (3 % of the layer FLOPs), far below
// a 128-row tcgen05 tile, so these kernels use warp-level mma.sync (HMMA) tiles.
#include "../../include/b200u.h"
#include "common.cuh"

namespace b200u {

constexpr int HD = 64;        // head dim (config/uniter-{base,large}.json: H / heads == 64)
constexpr int SROW = HD + 8;  // padded smem row (144 B) -> conflict-free ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
        "{%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// A-operand fragments (16 rows x 64 cols, row-major smem tile with stride SROW) -> 4 k-steps.
__device__ __forceinline__ void load_a_frags(const bf16* tile, int lane, uint32_t (&a)[4][4]) {
    const int r = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int c = (lane >> 4) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
        ldsm_x4(smem_u32(tile + r * SROW + ks * 16 + c), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}

// acc[2 n-tiles] += A(16x64) · X[n0..n0+15, 0..63]ᵀ      (X row-major: "B[k][n] = X[n][k]")
__device__ __forceinline__ void mma_xt(float (&c0)[4], float (&c1)[4], const uint32_t (&a)[4][4],
                                       const bf16* X, int n0, int lane) {
    const int r = n0 + (lane & 7) + (lane >> 4) * 8;
    const int c = ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(X + r * SROW + ks * 16 + c), b0, b1, b2, b3);
        mma16816(c0, a[ks][0], a[ks][1], a[ks][2], a[ks][3], b0, b1);
        mma16816(c1, a[ks][0], a[ks][1], a[ks][2], a[ks][3], b2, b3);
    }
}

// out[8 n-tiles over 64 cols] += P(16 x 16, k-step at rows k0..k0+15 of X) · X[k0.., 0..63]
__device__ __forceinline__ void mma_x(float (&o)[8][4], uint32_t a0, uint32_t a1, uint32_t a2,
                                      uint32_t a3, const bf16* X, int k0, int lane) {
    const int r = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int c = (lane >> 4) * 8;
#pragma unroll
    for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(X + r * SROW + np * 16 + c), b0, b1, b2, b3);
        mma16816(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma16816(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
}

// cooperative load of `rows` rows x 64 bf16 from a [*, ld] matrix into a padded smem tile,
// zero-filling rows >= valid.
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, int ld, int valid, int rows,
                                          int tid, int nthreads) {
    for (int i = tid; i < rows * 8; i += nthreads) {
        const int r = i >> 3, ch = i & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < valid) v = *reinterpret_cast<const uint4*>(src + (size_t)r * ld + ch * 8);
        *reinterpret_cast<uint4*>(dst + r * SROW + ch * 8) = v;
    }
}

__device__ __forceinline__ uint32_t attn_pair_base(int bh, int i, int L) {
    return (uint32_t)((bh * L + i) * ((L + 1) >> 1));
}

constexpr float LOG2E = 1.4426950408889634f;

// ---------------------------------------------------------------------------------------
// Forward. grid = (B*heads, ceil(L/64)), 4 warps, warp w owns query rows q0 + 16w .. +15.
// ---------------------------------------------------------------------------------------
template <int LP>
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ mask, bf16* __restrict__ ctx,
                float* __restrict__ lse, int L, int nh, int H, DropoutCfg drop) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* sK = reinterpret_cast<bf16*>(smem_raw);
    bf16* sV = sK + LP * SROW;
    bf16* sQ = sV + LP * SROW;
    float* sM = reinterpret_cast<float*>(sQ + 64 * SROW);

    const int bh = blockIdx.x, b = bh / nh, h = bh - b * nh;
    const int q0 = blockIdx.y * 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = 3 * H;
    const bf16* base = qkv + (size_t)b * L * ld + h * HD;

    load_tile(sK, base + H, ld, L, LP, tid, 128);
    load_tile(sV, base + 2 * H, ld, L, LP, tid, 128);
    load_tile(sQ, base + (size_t)q0 * ld, ld, L - q0, 64, tid, 128);
    for (int j = tid; j < LP; j += 128)
        sM[j] = (j < L) ? mask[(size_t)b * L + j] : -INFINITY;
    __syncthreads();

    const int r0 = q0 + warp * 16;  // first query row of this warp
    if (r0 >= L) return;

    uint32_t qa[4][4];
    load_a_frags(sQ + warp * 16 * SROW, lane, qa);

    constexpr int NT = LP / 8;
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int np = 0; np < NT / 2; ++np)
        if (np * 16 < L) mma_xt(s[2 * np], s[2 * np + 1], qa, sK, np * 16, lane);

    // scores / sqrt(64) + additive mask, row max (rows g and g+8 of the strip)
    const int t2 = (lane & 3) * 2;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const float m0 = sM[nt * 8 + t2], m1 = sM[nt * 8 + t2 + 1];
        s[nt][0] = s[nt][0] * 0.125f + m0;
        s[nt][1] = s[nt][1] * 0.125f + m1;
        s[nt][2] = s[nt][2] * 0.125f + m0;
        s[nt][3] = s[nt][3] * 0.125f + m1;
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = exp2f((s[nt][0] - mx0) * LOG2E);
        s[nt][1] = exp2f((s[nt][1] - mx0) * LOG2E);
        s[nt][2] = exp2f((s[nt][2] - mx1) * LOG2E);
        s[nt][3] = exp2f((s[nt][3] - mx1) * LOG2E);
        sum0 += s[nt][0] + s[nt][1];
        sum1 += s[nt][2] + s[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    const int g = lane >> 2;
    const int i0 = r0 + g, i1 = r0 + g + 8;
    if ((lane & 3) == 0 && lse) {
        if (i0 < L) lse[(size_t)bh * L + i0] = mx0 + logf(sum0);
        if (i1 < L) lse[(size_t)bh * L + i1] = mx1 + logf(sum1);
    }
    float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;

    // dropout on the probabilities (model/layer.py:95) folded into the normalisation
    const uint64_t seed = load_seed(drop);
    const uint32_t pb0 = attn_pair_base(bh, i0, L), pb1 = attn_pair_base(bh, i1, L);
    if (drop.thresh16) { inv0 *= drop.scale; inv1 *= drop.scale; }

    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < NT / 2; ++kt) {
        if (kt * 16 < L) {
            float p[2][4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int nt = 2 * kt + e;
                p[e][0] = s[nt][0] * inv0; p[e][1] = s[nt][1] * inv0;
                p[e][2] = s[nt][2] * inv1; p[e][3] = s[nt][3] * inv1;
                if (drop.thresh16) {
                    const uint32_t jp = (uint32_t)(nt * 8 + t2) >> 1;
                    const uint32_t h0 = rng_pair(seed, drop.stream, pb0 + jp);
                    const uint32_t h1 = rng_pair(seed, drop.stream, pb1 + jp);
                    if ((h0 & 0xffffu) < drop.thresh16) p[e][0] = 0.f;
                    if ((h0 >> 16) < drop.thresh16) p[e][1] = 0.f;
                    if ((h1 & 0xffffu) < drop.thresh16) p[e][2] = 0.f;
                    if ((h1 >> 16) < drop.thresh16) p[e][3] = 0.f;
                }
            }
            mma_x(o, pack_bf16(p[0][0], p[0][1]), pack_bf16(p[0][2], p[0][3]),
                  pack_bf16(p[1][0], p[1][1]), pack_bf16(p[1][2], p[1][3]), sV, kt * 16, lane);
        }
    }

    bf16* out = ctx + (size_t)b * L * H + h * HD;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        if (i0 < L) *reinterpret_cast<uint32_t*>(out + (size_t)i0 * H + nt * 8 + t2) = pack_bf16(o[nt][0], o[nt][1]);
        if (i1 < L) *reinterpret_cast<uint32_t*>(out + (size_t)i1 * H + nt * 8 + t2) = pack_bf16(o[nt][2], o[nt][3]);
    }
}

// ---------------------------------------------------------------------------------------
// Backward. grid = (B*heads, 2): blockIdx.y == 0 computes dQ (warps own 16 query rows and
// sweep the keys in chunks of 64), blockIdx.y == 1 computes dK and dV (warps own 16 key rows
// and sweep the queries), both from recomputed probabilities. No atomics, deterministic.
//   D_i   = sum_d dO[i,d] * O[i,d]
//   dPd   = dO·Vᵀ ; dP = dropmask*scale*dPd ; dS = P*(dP - D_i) ; dscore = dS / 8
//   dQ = dscore·K ; dK = dscoreᵀ·Q ; dV = Pdᵀ·dO
// ---------------------------------------------------------------------------------------
template <int LP>
__global__ void __launch_bounds__(256)
attn_bwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ mask,
                const bf16* __restrict__ ctx, const bf16* __restrict__ dctx,
                const float* __restrict__ lse, bf16* __restrict__ dqkv, int L, int nh, int H,
                DropoutCfg drop) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
    bf16* sK = sQ + LP * SROW;
    bf16* sV = sK + LP * SROW;
    bf16* sdO = sV + LP * SROW;
    float* sM = reinterpret_cast<float*>(sdO + LP * SROW);
    float* sLse = sM + LP;
    float* sD = sLse + LP;

    const int bh = blockIdx.x, b = bh / nh, h = bh - b * nh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int ld = 3 * H;
    const bf16* base = qkv + (size_t)b * L * ld + h * HD;
    const bf16* dO = dctx + (size_t)b * L * H + h * HD;
    const bf16* O = ctx + (size_t)b * L * H + h * HD;

    load_tile(sQ, base, ld, L, LP, tid, blockDim.x);
    load_tile(sK, base + H, ld, L, LP, tid, blockDim.x);
    load_tile(sV, base + 2 * H, ld, L, LP, tid, blockDim.x);
    load_tile(sdO, dO, H, L, LP, tid, blockDim.x);
    for (int j = tid; j < LP; j += blockDim.x) {
        sM[j] = (j < L) ? mask[(size_t)b * L + j] : -INFINITY;
        sLse[j] = (j < L) ? lse[(size_t)bh * L + j] : 0.f;
    }
    __syncthreads();
    // D_i: 8 threads per row, 8 elements each
    for (int i = tid >> 3; i < LP; i += blockDim.x >> 3) {
        float d = 0.f;
        if (i < L) {
            const int ch = tid & 7;
            uint4 ov = *reinterpret_cast<const uint4*>(O + (size_t)i * H + ch * 8);
            uint4 dv = *reinterpret_cast<const uint4*>(sdO + i * SROW + ch * 8);
            const uint32_t* op = &ov.x;
            const uint32_t* dp = &dv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float2 a = unpack_bf16(op[k]), c = unpack_bf16(dp[k]);
                d += a.x * c.x + a.y * c.y;
            }
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        if ((tid & 7) == 0) sD[i] = d;
    }
    __syncthreads();

    const uint64_t seed = load_seed(drop);
    const int g = lane >> 2, t2 = (lane & 3) * 2;
    const int nblk = (L + 15) / 16;
    bf16* dbase = dqkv + (size_t)b * L * ld + h * HD;

    if (blockIdx.y == 0) {
        // ---------------- dQ: rows = queries ----------------
        for (int blk = warp; blk < nblk; blk += nwarps) {
            const int r0 = blk * 16;
            const int i0 = r0 + g, i1 = r0 + g + 8;
            uint32_t qa[4][4], da[4][4];
            load_a_frags(sQ + r0 * SROW, lane, qa);
            load_a_frags(sdO + r0 * SROW, lane, da);
            const float l0 = sLse[i0], l1 = sLse[i1];
            const float D0 = sD[i0], D1 = sD[i1];
            const uint32_t pb0 = attn_pair_base(bh, i0, L), pb1 = attn_pair_base(bh, i1, L);
            float dq[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
            for (int c0 = 0; c0 < L; c0 += 64) {
                float s[8][4], dp[8][4];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
                    dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
                }
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    if (c0 + np * 16 < L) {
                        mma_xt(s[2 * np], s[2 * np + 1], qa, sK, c0 + np * 16, lane);
                        mma_xt(dp[2 * np], dp[2 * np + 1], da, sV, c0 + np * 16, lane);
                    }
                }
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    if (c0 + np * 16 < L) {
                        float ds[2][4];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int nt = 2 * np + e;
                            const int j = c0 + nt * 8 + t2;
                            const float m0 = sM[j], m1 = sM[j + 1];
                            float p0 = exp2f((s[nt][0] * 0.125f + m0 - l0) * LOG2E);
                            float p1 = exp2f((s[nt][1] * 0.125f + m1 - l0) * LOG2E);
                            float p2 = exp2f((s[nt][2] * 0.125f + m0 - l1) * LOG2E);
                            float p3 = exp2f((s[nt][3] * 0.125f + m1 - l1) * LOG2E);
                            float d0 = dp[nt][0], d1 = dp[nt][1], d2 = dp[nt][2], d3 = dp[nt][3];
                            if (drop.thresh16) {
                                const uint32_t jp = (uint32_t)j >> 1;
                                const uint32_t h0 = rng_pair(seed, drop.stream, pb0 + jp);
                                const uint32_t h1 = rng_pair(seed, drop.stream, pb1 + jp);
                                d0 = ((h0 & 0xffffu) >= drop.thresh16) ? d0 * drop.scale : 0.f;
                                d1 = ((h0 >> 16) >= drop.thresh16) ? d1 * drop.scale : 0.f;
                                d2 = ((h1 & 0xffffu) >= drop.thresh16) ? d2 * drop.scale : 0.f;
                                d3 = ((h1 >> 16) >= drop.thresh16) ? d3 * drop.scale : 0.f;
                            }
                            ds[e][0] = p0 * (d0 - D0) * 0.125f;
                            ds[e][1] = p1 * (d1 - D0) * 0.125f;
                            ds[e][2] = p2 * (d2 - D1) * 0.125f;
                            ds[e][3] = p3 * (d3 - D1) * 0.125f;
                        }
                        mma_x(dq, pack_bf16(ds[0][0], ds[0][1]), pack_bf16(ds[0][2], ds[0][3]),
                              pack_bf16(ds[1][0], ds[1][1]), pack_bf16(ds[1][2], ds[1][3]), sK,
                              c0 + np * 16, lane);
                    }
                }
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                if (i0 < L) *reinterpret_cast<uint32_t*>(dbase + (size_t)i0 * ld + nt * 8 + t2) = pack_bf16(dq[nt][0], dq[nt][1]);
                if (i1 < L) *reinterpret_cast<uint32_t*>(dbase + (size_t)i1 * ld + nt * 8 + t2) = pack_bf16(dq[nt][2], dq[nt][3]);
            }
        }
    } else {
        // ---------------- dK, dV: rows = keys, columns = queries ----------------
        const int Lh = (L + 1) >> 1;
        for (int blk = warp; blk < nblk; blk += nwarps) {
            const int r0 = blk * 16;
            const int j0 = r0 + g, j1 = r0 + g + 8;
            uint32_t ka[4][4], va[4][4];
            load_a_frags(sK + r0 * SROW, lane, ka);
            load_a_frags(sV + r0 * SROW, lane, va);
            const float m0 = sM[j0], m1 = sM[j1];
            float dk[8][4], dv[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
                dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
            }
            for (int c0 = 0; c0 < L; c0 += 64) {
                float s[8][4], dp[8][4];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
                    dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
                }
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    if (c0 + np * 16 < L) {
                        mma_xt(s[2 * np], s[2 * np + 1], ka, sQ, c0 + np * 16, lane);     // Sᵀ
                        mma_xt(dp[2 * np], dp[2 * np + 1], va, sdO, c0 + np * 16, lane);  // dPdᵀ
                    }
                }
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    if (c0 + np * 16 < L) {
                        float pd[2][4], ds[2][4];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int nt = 2 * np + e;
                            const int i = c0 + nt * 8 + t2;  // query index of columns i, i+1
                            const float la = sLse[i], lb = sLse[i + 1];
                            const float Da = sD[i], Db = sD[i + 1];
                            float p0 = exp2f((s[nt][0] * 0.125f + m0 - la) * LOG2E);  // (j0, i)
                            float p1 = exp2f((s[nt][1] * 0.125f + m0 - lb) * LOG2E);  // (j0, i+1)
                            float p2 = exp2f((s[nt][2] * 0.125f + m1 - la) * LOG2E);  // (j1, i)
                            float p3 = exp2f((s[nt][3] * 0.125f + m1 - lb) * LOG2E);  // (j1, i+1)
                            // queries beyond L have zero dO and zero Q rows; their lse is 0 -> clamp p
                            if (i >= L) { p0 = 0.f; p2 = 0.f; }
                            if (i + 1 >= L) { p1 = 0.f; p3 = 0.f; }
                            float k0 = 1.f, k1 = 1.f, k2 = 1.f, k3 = 1.f;
                            if (drop.thresh16) {
                                const uint32_t ba = (uint32_t)((bh * L + i) * Lh);
                                const uint32_t bb = ba + (uint32_t)Lh;
                                const uint32_t ha0 = rng_pair(seed, drop.stream, ba + ((uint32_t)j0 >> 1));
                                const uint32_t hb0 = rng_pair(seed, drop.stream, bb + ((uint32_t)j0 >> 1));
                                const uint32_t ha1 = rng_pair(seed, drop.stream, ba + ((uint32_t)j1 >> 1));
                                const uint32_t hb1 = rng_pair(seed, drop.stream, bb + ((uint32_t)j1 >> 1));
                                const int sh0 = (j0 & 1) * 16, sh1 = (j1 & 1) * 16;
                                k0 = (((ha0 >> sh0) & 0xffffu) >= drop.thresh16) ? drop.scale : 0.f;
                                k1 = (((hb0 >> sh0) & 0xffffu) >= drop.thresh16) ? drop.scale : 0.f;
                                k2 = (((ha1 >> sh1) & 0xffffu) >= drop.thresh16) ? drop.scale : 0.f;
                                k3 = (((hb1 >> sh1) & 0xffffu) >= drop.thresh16) ? drop.scale : 0.f;
                            }
                            pd[e][0] = p0 * k0; pd[e][1] = p1 * k1; pd[e][2] = p2 * k2; pd[e][3] = p3 * k3;
                            ds[e][0] = p0 * (dp[nt][0] * k0 - Da) * 0.125f;
                            ds[e][1] = p1 * (dp[nt][1] * k1 - Db) * 0.125f;
                            ds[e][2] = p2 * (dp[nt][2] * k2 - Da) * 0.125f;
                            ds[e][3] = p3 * (dp[nt][3] * k3 - Db) * 0.125f;
                        }
                        mma_x(dv, pack_bf16(pd[0][0], pd[0][1]), pack_bf16(pd[0][2], pd[0][3]),
                              pack_bf16(pd[1][0], pd[1][1]), pack_bf16(pd[1][2], pd[1][3]), sdO,
                              c0 + np * 16, lane);
                        mma_x(dk, pack_bf16(ds[0][0], ds[0][1]), pack_bf16(ds[0][2], ds[0][3]),
                              pack_bf16(ds[1][0], ds[1][1]), pack_bf16(ds[1][2], ds[1][3]), sQ,
                              c0 + np * 16, lane);
                    }
                }
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                if (j0 < L) {
                    *reinterpret_cast<uint32_t*>(dbase + (size_t)j0 * ld + H + nt * 8 + t2) = pack_bf16(dk[nt][0], dk[nt][1]);
                    *reinterpret_cast<uint32_t*>(dbase + (size_t)j0 * ld + 2 * H + nt * 8 + t2) = pack_bf16(dv[nt][0], dv[nt][1]);
                }
                if (j1 < L) {
                    *reinterpret_cast<uint32_t*>(dbase + (size_t)j1 * ld + H + nt * 8 + t2) = pack_bf16(dk[nt][2], dk[nt][3]);
                    *reinterpret_cast<uint32_t*>(dbase + (size_t)j1 * ld + 2 * H + nt * 8 + t2) = pack_bf16(dv[nt][2], dv[nt][3]);
                }
            }
        }
    }
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

template <int LP>
static int launch_fwd(const void* qkv, const float* mask, void* ctx, float* lse, int B, int L, int nh,
                      int H, DropoutCfg dc, cudaStream_t stream) {
    const size_t smem = (size_t)(2 * LP + 64) * SROW * 2 + LP * 4;
    static bool set = false;
    if (!set) {
        B200U_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<LP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        set = true;
    }
    attn_fwd_kernel<LP><<<dim3(B * nh, (L + 63) / 64), 128, smem, stream>>>((const bf16*)qkv, mask, (bf16*)ctx, lse, L, nh, H, dc);
    B200U_CHECK_LAUNCH("attn_fwd_kernel");
    return B200U_OK;
}

template <int LP>
static int launch_bwd(const void* qkv, const float* mask, const void* ctx, const void* dctx,
                      const float* lse, void* dqkv, int B, int L, int nh, int H, DropoutCfg dc,
                      cudaStream_t stream) {
    const size_t smem = (size_t)4 * LP * SROW * 2 + 3 * LP * 4;
    static bool set = false;
    if (!set) {
        B200U_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<LP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        set = true;
    }
    const int nblk = (L + 15) / 16;
    int warps = nblk < 4 ? 4 : (nblk > 8 ? 8 : nblk);
    if (nblk > 8) warps = (nblk + 1) / 2 > 8 ? 8 : (nblk + 1) / 2;  // two balanced rounds
    attn_bwd_kernel<LP><<<dim3(B * nh, 2), warps * 32, smem, stream>>>((const bf16*)qkv, mask, (const bf16*)ctx, (const bf16*)dctx, lse, (bf16*)dqkv, L, nh, H, dc);
    B200U_CHECK_LAUNCH("attn_bwd_kernel");
    return B200U_OK;
}

}  // namespace b200u

using namespace b200u;

extern "C" int b200u_attention_fwd(const void* qkv, const float* mask, void* ctx, float* lse, int B,
                                   int L, int num_heads, int H, const b200u_dropout_t* drop,
                                   b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(qkv && mask && ctx, "attention_fwd: null pointer");
    B200U_CHECK_ARG(num_heads > 0 && H == num_heads * HD, "attention_fwd: head dim must be 64 (H=%d heads=%d)", H, num_heads);
    B200U_CHECK_ARG(L > 0 && L <= 256, "attention_fwd: joint sequence length %d unsupported (1..256)", L);
    if (B == 0) return B200U_OK;
    DropoutCfg dc = make_drop(drop);
    B200U_CHECK_ARG(dc.thresh16 == 0 || dc.seed_ptr, "attention_fwd: dropout needs seed_ptr");
    if (L <= 64) return launch_fwd<64>(qkv, mask, ctx, lse, B, L, num_heads, H, dc, stream);
    if (L <= 128) return launch_fwd<128>(qkv, mask, ctx, lse, B, L, num_heads, H, dc, stream);
    if (L <= 192) return launch_fwd<192>(qkv, mask, ctx, lse, B, L, num_heads, H, dc, stream);
    return launch_fwd<256>(qkv, mask, ctx, lse, B, L, num_heads, H, dc, stream);
}

extern "C" int b200u_attention_bwd(const void* qkv, const float* mask, const void* ctx,
                                   const void* dctx, const float* lse, void* dqkv, int B, int L,
                                   int num_heads, int H, const b200u_dropout_t* drop,
                                   b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(qkv && mask && ctx && dctx && lse && dqkv, "attention_bwd: null pointer");
    B200U_CHECK_ARG(num_heads > 0 && H == num_heads * HD, "attention_bwd: head dim must be 64 (H=%d heads=%d)", H, num_heads);
    B200U_CHECK_ARG(L > 0 && L <= 256, "attention_bwd: joint sequence length %d unsupported (1..256)", L);
    if (B == 0) return B200U_OK;
    DropoutCfg dc = make_drop(drop);
    B200U_CHECK_ARG(dc.thresh16 == 0 || dc.seed_ptr, "attention_bwd: dropout needs seed_ptr");
    if (L <= 64) return launch_bwd<64>(qkv, mask, ctx, dctx, lse, dqkv, B, L, num_heads, H, dc, stream);
    if (L <= 128) return launch_bwd<128>(qkv, mask, ctx, dctx, lse, dqkv, B, L, num_heads, H, dc, stream);
    if (L <= 192) return launch_bwd<192>(qkv, mask, ctx, dctx, lse, dqkv, B, L, num_heads, H, dc, stream);
    return launch_bwd<256>(qkv, mask, ctx, dctx, lse, dqkv, B, L, num_heads, H, dc, stream);
}
