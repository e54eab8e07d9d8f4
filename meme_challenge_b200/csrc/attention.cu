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
// Both kernels work on 64-row tiles (576 / 1152 CTAs at the C2 shape) and are sized for 3-4
// co-resident CTAs per SM, which is what hides the ldmatrix / MUFU latencies.
//
// The joint sequence is at most 164 long with 64-wide heads (3 % of the layer FLOPs), far below
// a 128-row tcgen05 tile, so these kernels use warp-level mma.sync (HMMA) tiles.
#include "../../include/b200u.h"
#include "common.cuh"

#include <stdlib.h>

namespace b200u {

// attention_tc.cu: tcgen05 / TMEM / TMA forward
int attention_fwd_tc(const void* qkv, const float* mask, void* ctx, float* lse, int B, int L, int nh, int H,
                     const b200u_dropout_t* drop, cudaStream_t stream);
int attention_bwd_tc(const void* qkv, const float* mask, const void* ctx, const void* dctx, const float* lse,
                     void* dqkv, float* dbias, int B, int L, int nh, int H, const b200u_dropout_t* drop,
                     cudaStream_t stream);
// b200u_set_attention_impl(): 1 (default) = tcgen05 kernels, 0 = the mma.sync kernels of this file
// (B200U_ATTN_TC=0 selects them at load time)
static int g_attn_tc = [] {
    const char* e = getenv("B200U_ATTN_TC");
    return (e && e[0] == '0') ? 0 : 1;
}();

constexpr int HD = 64;  // head dim (config/uniter-{base,large}.json: H / heads == 64)
constexpr int TQ = 64;  // rows per CTA tile (4 warps x 16)

// smem tiles are [rows][64] bf16 with 128-byte rows and the 16-byte chunk index XOR-swizzled by
// (row & 7): dense (no padding) and conflict-free for ldmatrix.
__device__ __forceinline__ int sw_off(int r, int chunk) { return r * HD + ((chunk ^ (r & 7)) << 3); }

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

// A-operand fragments: rows row0..row0+15 of a swizzled tile, all 64 columns -> 4 k-steps.
__device__ __forceinline__ void load_a_frags(const bf16* tile, int row0, int lane, uint32_t (&a)[4][4]) {
    const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int c = lane >> 4;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
        ldsm_x4(smem_u32(tile + sw_off(r, ks * 2 + c)), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}

// acc[2 n-tiles] += A(16x64) · X[n0..n0+15, 0..63]ᵀ      (X row-major: "B[k][n] = X[n][k]")
__device__ __forceinline__ void mma_xt(float (&c0)[4], float (&c1)[4], const uint32_t (&a)[4][4],
                                       const bf16* X, int n0, int lane) {
    const int r = n0 + (lane & 7) + (lane >> 4) * 8;
    const int c = (lane >> 3) & 1;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(X + sw_off(r, ks * 2 + c)), b0, b1, b2, b3);
        mma16816(c0, a[ks][0], a[ks][1], a[ks][2], a[ks][3], b0, b1);
        mma16816(c1, a[ks][0], a[ks][1], a[ks][2], a[ks][3], b2, b3);
    }
}

// out[8 n-tiles over 64 cols] += P(16 x 16, k-step at rows k0..k0+15 of X) · X[k0.., 0..63]
__device__ __forceinline__ void mma_x(float (&o)[8][4], uint32_t a0, uint32_t a1, uint32_t a2,
                                      uint32_t a3, const bf16* X, int k0, int lane) {
    const int r = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int c = lane >> 4;
#pragma unroll
    for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(X + sw_off(r, np * 2 + c)), b0, b1, b2, b3);
        mma16816(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma16816(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
}

// cooperative ASYNC load (cp.async, 16 B per request, all requests in flight at once) of `rows` rows
// x 64 bf16 from a [*, ld] matrix into a swizzled smem tile, zero-filling rows >= valid. Callers
// finish with load_tiles_wait() before the __syncthreads() that publishes the tiles.
__device__ __forceinline__ void load_rows(bf16* dst, const bf16* src, int ld, int valid, int rb, int re,
                                          int tid, int nthreads, int valid_chunks = 8) {
    // rows [rb, re) of the tile. A thread keeps its 16-byte chunk column and walks the rows in steps
    // of nthreads/8, so the source stride and the chunk predicate are loop invariant (and, when the
    // step is a multiple of 8 rows, so is the swizzled chunk).
    const int ch = tid & 7;
    const int rstep = nthreads >> 3;
    int r = rb + (tid >> 3);
    const bf16* s = src + (size_t)r * ld + ch * 8;
    const size_t sstep = (size_t)rstep * ld;
    const bool chunk_ok = ch < valid_chunks;
    for (; r < re; r += rstep, s += sstep) {
        bf16* d = dst + r * HD + ((ch ^ (r & 7)) << 3);
        if (chunk_ok && r < valid) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d)), "l"(s) : "memory");
        } else {
            *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
        }
    }
}
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, int ld, int valid, int rows,
                                          int tid, int nthreads, int valid_chunks = 8) {
    load_rows(dst, src, ld, valid, 0, rows, tid, nthreads, valid_chunks);
}
__device__ __forceinline__ void load_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `pending` of this thread's committed cp.async groups are still in flight
__device__ __forceinline__ void load_wait_pending(int pending) {
    if (pending <= 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else if (pending == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else if (pending == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else asm volatile("cp.async.wait_group 3;" ::: "memory");
}
__device__ __forceinline__ void load_tiles_wait() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

constexpr float LOG2E = 1.4426950408889634f;

// ---------------------------------------------------------------------------------------
// Forward. grid = (B*heads, ceil(L/64)); 4 warps, warp w owns query rows q0+16w..+15 and sweeps
// the keys in chunks of 64 with an online softmax, so the live state is 32 score + 32 output
// registers per thread and four CTAs share an SM.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 4)
attn_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ mask, bf16* __restrict__ ctx,
                float* __restrict__ lse, int L, int LP, int nh, int H, DropoutCfg drop) {
    pdl_sync();
    extern __shared__ __align__(128) uint8_t smem_raw[];
    bf16* sK = reinterpret_cast<bf16*>(smem_raw);
    bf16* sV = sK + LP * HD;
    bf16* sQ = sV + LP * HD;
    float* sM = reinterpret_cast<float*>(sQ + TQ * HD);

    const int bh = blockIdx.x, b = bh / nh, h = bh - b * nh;
    const int q0 = blockIdx.y * TQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = 3 * H;
    const bf16* base = qkv + (size_t)b * L * ld + h * HD;

    // K/V arrive in 64-key chunks, one cp.async group each (the first also carries Q): the sweep
    // below starts on chunk 0 while the later chunks are still in flight
    const int nchunks = (L + 63) >> 6;
    load_tile(sQ, base + (size_t)q0 * ld, ld, L - q0, TQ, tid, 128);
    for (int c = 0; c < nchunks; ++c) {
        const int rb = c * 64, re = min(LP, rb + 64);
        load_rows(sK, base + H, ld, L, rb, re, tid, 128);
        load_rows(sV, base + 2 * H, ld, L, rb, re, tid, 128);
        load_commit();
    }
    // additive mask in the log2 domain, padded with -inf to the 64-key chunks the loop sweeps
    const int LP64 = (L + 63) & ~63;
    for (int j = tid; j < LP64; j += 128) sM[j] = (j < L) ? mask[(size_t)b * L + j] * LOG2E : -INFINITY;

    const int r0 = q0 + warp * 16;
    const bool active = r0 < L;  // warps past the end of the sequence only take part in the barriers
    uint32_t qa[4][4];

    const int g = lane >> 2, t2 = (lane & 3) * 2;
    const int i0 = r0 + g, i1 = r0 + g + 8;
    const uint32_t key = drop.thresh16 ? attn_key(load_seed(drop), drop.stream) : 0u;
    const uint32_t pb0 = attn_block_base(bh, i0, L), pb1 = attn_block_base(bh, i1, L);
    constexpr float SC = 0.125f * LOG2E;  // scores / sqrt(64) (model/layer.py:86), in log2 units

    // running max m and sum l of exp2(score2 - m), score2 = (q.k / 8 + mask) * log2(e)
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;

    for (int c0 = 0; c0 < L; c0 += 64) {
        load_wait_pending(nchunks - 1 - (c0 >> 6));
        __syncthreads();
        if (!active) continue;
        if (c0 == 0) load_a_frags(sQ, warp * 16, lane, qa);
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int np = 0; np < 4; ++np)
            if (c0 + np * 16 < L) mma_xt(s[2 * np], s[2 * np + 1], qa, sK, c0 + np * 16, lane);
        // scores / sqrt(64) + additive mask (model/layer.py:86-88), chunk row-max
        float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float2 mk = *reinterpret_cast<const float2*>(sM + c0 + nt * 8 + t2);
            s[nt][0] = fmaf(s[nt][0], SC, mk.x); s[nt][1] = fmaf(s[nt][1], SC, mk.y);
            s[nt][2] = fmaf(s[nt][2], SC, mk.x); s[nt][3] = fmaf(s[nt][3], SC, mk.y);
            cm0 = fmaxf(cm0, fmaxf(s[nt][0], s[nt][1]));
            cm1 = fmaxf(cm1, fmaxf(s[nt][2], s[nt][3]));
        }
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
        const float n0 = fmaxf(m0, cm0), n1 = fmaxf(m1, cm1);
        const float corr0 = ex2_approx(m0 - n0), corr1 = ex2_approx(m1 - n1);
        m0 = n0; m1 = n1;
        l0 *= corr0; l1 *= corr1;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            o[nt][0] *= corr0; o[nt][1] *= corr0; o[nt][2] *= corr1; o[nt][3] *= corr1;
        }
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            if (c0 + kt * 16 < L) {
                float p[2][4];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int nt = 2 * kt + e;
                    p[e][0] = ex2_approx(s[nt][0] - m0); p[e][1] = ex2_approx(s[nt][1] - m0);
                    p[e][2] = ex2_approx(s[nt][2] - m1); p[e][3] = ex2_approx(s[nt][3] - m1);
                    l0 += p[e][0] + p[e][1];
                    l1 += p[e][2] + p[e][3];
                    if (drop.thresh16) {  // dropout on the probabilities (model/layer.py:95)
                        const uint32_t jp = (uint32_t)(c0 + nt * 8 + t2) >> 1;
                        const uint32_t h0 = attn_rng(key, pb0 + jp, i0 & 1), h1 = attn_rng(key, pb1 + jp, i1 & 1);
                        if ((h0 & 0xffffu) < drop.thresh16) p[e][0] = 0.f;
                        if ((h0 >> 16) < drop.thresh16) p[e][1] = 0.f;
                        if ((h1 & 0xffffu) < drop.thresh16) p[e][2] = 0.f;
                        if ((h1 >> 16) < drop.thresh16) p[e][3] = 0.f;
                    }
                }
                mma_x(o, pack_bf16(p[0][0], p[0][1]), pack_bf16(p[0][2], p[0][3]),
                      pack_bf16(p[1][0], p[1][1]), pack_bf16(p[1][2], p[1][3]), sV, c0 + kt * 16, lane);
            }
        }
    }
    if (!active) return;
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    if ((lane & 3) == 0 && lse) {
        // natural-log log-sum-exp of the masked, scaled scores
        if (i0 < L) lse[(size_t)bh * L + i0] = (m0 + log2f(l0)) * 0.69314718055994530942f;
        if (i1 < L) lse[(size_t)bh * L + i1] = (m1 + log2f(l1)) * 0.69314718055994530942f;
    }
    const float inv0 = (drop.thresh16 ? drop.scale : 1.0f) / l0, inv1 = (drop.thresh16 ? drop.scale : 1.0f) / l1;
    bf16* out = ctx + (size_t)b * L * H + h * HD;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        if (i0 < L) *reinterpret_cast<uint32_t*>(out + (size_t)i0 * H + nt * 8 + t2) = pack_bf16(o[nt][0] * inv0, o[nt][1] * inv0);
        if (i1 < L) *reinterpret_cast<uint32_t*>(out + (size_t)i1 * H + nt * 8 + t2) = pack_bf16(o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

// ---------------------------------------------------------------------------------------
// Backward, two launches (the second consumes what the first leaves in L2):
//  1. attn_bwd_dq_kernel, grid (B*heads, ceil(L/64)): a 64-row query tile against all keys in
//     chunks of 32. Recomputes P from the saved log-sum-exp, regenerates the dropout mask, forms
//         D_i = sum_d dO[i,d] O[i,d] ; dPd = dO·Vᵀ ; dP = dropmask*scale*dPd ; dS = P*(dP - D_i)/8
//     accumulates dQ = dS·K and ALSO writes Pd = dropmask*scale*P and dS (bf16, [B*heads, L, LP])
//     to a scratch buffer, so the element-wise work (exp2, RNG) is done once.
//  2. attn_bwd_dkv_kernel, grid (B*heads, ceil(L/64)): a 64-row key tile; dV = Pdᵀ·dO and
//     dK = dSᵀ·Q are plain tensor-core products over the scratch tiles (ldmatrix.trans A operands).
// No atomics, deterministic.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 3)
attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const float* __restrict__ mask,
                   const bf16* __restrict__ ctx, const bf16* __restrict__ dctx,
                   const float* __restrict__ lse, bf16* __restrict__ dqkv, bf16* __restrict__ scrP,
                   bf16* __restrict__ scrS, int L, int LP, int nh, int H, DropoutCfg drop) {
    pdl_sync();
    extern __shared__ __align__(128) uint8_t smem_raw[];
    bf16* sK = reinterpret_cast<bf16*>(smem_raw);
    bf16* sV = sK + LP * HD;
    bf16* sQ = sV + LP * HD;
    bf16* sdO = sQ + TQ * HD;
    float* sM2 = reinterpret_cast<float*>(sdO + TQ * HD);  // additive mask * log2(e)
    float* sL2 = sM2 + LP;                                 // lse * log2(e) of the tile rows
    float* sD = sL2 + TQ;                                  // D of the tile rows

    const int bh = blockIdx.x, b = bh / nh, h = bh - b * nh;
    const int t0 = blockIdx.y * TQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = 3 * H;
    const bf16* base = qkv + (size_t)b * L * ld + h * HD;
    const bf16* dO = dctx + (size_t)b * L * H + h * HD;
    const bf16* O = ctx + (size_t)b * L * H + h * HD;

    // Q/dO tile + K/V in 64-key chunks, one cp.async group per chunk (the first carries Q and dO)
    const int nchunks = (LP + 63) >> 6;
    load_tile(sQ, base + (size_t)t0 * ld, ld, L - t0, TQ, tid, 128);
    load_tile(sdO, dO + (size_t)t0 * H, H, L - t0, TQ, tid, 128);
    for (int c = 0; c < nchunks; ++c) {
        const int rb = c * 64, re = min(LP, rb + 64);
        load_rows(sK, base + H, ld, L, rb, re, tid, 128);
        load_rows(sV, base + 2 * H, ld, L, rb, re, tid, 128);
        load_commit();
    }
    for (int j = tid; j < LP; j += 128) sM2[j] = (j < L) ? mask[(size_t)b * L + j] * LOG2E : -INFINITY;
    for (int i = tid >> 3; i < TQ; i += 16) {  // D_i and lse_i of the tile rows: 8 threads per row
        const int gi = t0 + i;
        float d = 0.f;
        if (gi < L) {
            const int ch = tid & 7;
            uint4 ov = *reinterpret_cast<const uint4*>(O + (size_t)gi * H + ch * 8);
            uint4 dv = *reinterpret_cast<const uint4*>(dO + (size_t)gi * H + ch * 8);
            const uint32_t* op = &ov.x;
            const uint32_t* dp = &dv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float2 x = unpack_bf16(op[k]), y = unpack_bf16(dp[k]);
                d += x.x * y.x + x.y * y.y;
            }
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        if ((tid & 7) == 0) {
            sD[i] = d;
            sL2[i] = (gi < L) ? lse[(size_t)bh * L + gi] * LOG2E : 0.f;
        }
    }
    const int r0 = t0 + warp * 16;
    const bool active = r0 < L;  // warps past the end of the sequence only take part in the barriers
    const uint32_t key = drop.thresh16 ? attn_key(load_seed(drop), drop.stream) : 0u;
    const int g = lane >> 2, t2 = (lane & 3) * 2;
    const int x0 = r0 + g, x1 = r0 + g + 8;
    uint32_t qa[4][4], da[4][4];
    float la = 0.f, lb = 0.f, Da = 0.f, Db = 0.f;
    const uint32_t pb0 = attn_block_base(bh, x0, L), pb1 = attn_block_base(bh, x1, L);
    constexpr float SC = 0.125f * LOG2E;
    bf16* pP0 = scrP + ((size_t)bh * L + x0) * LP;
    bf16* pP1 = scrP + ((size_t)bh * L + x1) * LP;
    bf16* pS0 = scrS + ((size_t)bh * L + x0) * LP;
    bf16* pS1 = scrS + ((size_t)bh * L + x1) * LP;
    float dq[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
    for (int c0 = 0; c0 < LP; c0 += 32) {
        if ((c0 & 63) == 0) {
            load_wait_pending(nchunks - 1 - (c0 >> 6));
            __syncthreads();
            if (c0 == 0 && active) {
                load_a_frags(sQ, warp * 16, lane, qa);
                load_a_frags(sdO, warp * 16, lane, da);
                la = sL2[warp * 16 + g]; lb = sL2[warp * 16 + g + 8];
                Da = sD[warp * 16 + g]; Db = sD[warp * 16 + g + 8];
            }
        }
        if (!active) continue;
        float s[4][4], dp[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
        }
#pragma unroll
        for (int np = 0; np < 2; ++np) {
            if (c0 + np * 16 < LP) {
                mma_xt(s[2 * np], s[2 * np + 1], qa, sK, c0 + np * 16, lane);    // Q·Kᵀ
                mma_xt(dp[2 * np], dp[2 * np + 1], da, sV, c0 + np * 16, lane);  // dO·Vᵀ
            }
        }
#pragma unroll
        for (int np = 0; np < 2; ++np) {
            if (c0 + np * 16 < LP) {
                uint32_t dsp[2][2], pdp[2][2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int nt = 2 * np + e;
                    const int j = c0 + nt * 8 + t2;
                    const float ma = sM2[j], mb = sM2[j + 1];
                    const float p0 = ex2_approx(fmaf(s[nt][0], SC, ma) - la);
                    const float p1 = ex2_approx(fmaf(s[nt][1], SC, mb) - la);
                    const float p2 = ex2_approx(fmaf(s[nt][2], SC, ma) - lb);
                    const float p3 = ex2_approx(fmaf(s[nt][3], SC, mb) - lb);
                    float k0 = 1.f, k1 = 1.f, k2 = 1.f, k3 = 1.f;
                    if (drop.thresh16) {
                        const uint32_t jp = (uint32_t)j >> 1;
                        const uint32_t h0 = attn_rng(key, pb0 + jp, x0 & 1), h1 = attn_rng(key, pb1 + jp, x1 & 1);
                        k0 = ((h0 & 0xffffu) >= drop.thresh16) ? drop.scale : 0.f;
                        k1 = ((h0 >> 16) >= drop.thresh16) ? drop.scale : 0.f;
                        k2 = ((h1 & 0xffffu) >= drop.thresh16) ? drop.scale : 0.f;
                        k3 = ((h1 >> 16) >= drop.thresh16) ? drop.scale : 0.f;
                    }
                    dsp[e][0] = pack_bf16(p0 * (dp[nt][0] * k0 - Da) * 0.125f, p1 * (dp[nt][1] * k1 - Da) * 0.125f);
                    dsp[e][1] = pack_bf16(p2 * (dp[nt][2] * k2 - Db) * 0.125f, p3 * (dp[nt][3] * k3 - Db) * 0.125f);
                    pdp[e][0] = pack_bf16(p0 * k0, p1 * k1);
                    pdp[e][1] = pack_bf16(p2 * k2, p3 * k3);
                    if (x0 < L) {
                        *reinterpret_cast<uint32_t*>(pP0 + j) = pdp[e][0];
                        *reinterpret_cast<uint32_t*>(pS0 + j) = dsp[e][0];
                    }
                    if (x1 < L) {
                        *reinterpret_cast<uint32_t*>(pP1 + j) = pdp[e][1];
                        *reinterpret_cast<uint32_t*>(pS1 + j) = dsp[e][1];
                    }
                }
                mma_x(dq, dsp[0][0], dsp[0][1], dsp[1][0], dsp[1][1], sK, c0 + np * 16, lane);
            }
        }
    }
    if (!active) return;
    bf16* dbase = dqkv + (size_t)b * L * ld + h * HD;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        if (x0 < L) *reinterpret_cast<uint32_t*>(dbase + (size_t)x0 * ld + nt * 8 + t2) = pack_bf16(dq[nt][0], dq[nt][1]);
        if (x1 < L) *reinterpret_cast<uint32_t*>(dbase + (size_t)x1 * ld + nt * 8 + t2) = pack_bf16(dq[nt][2], dq[nt][3]);
    }
}

// acc[8 n-tiles] += Xᵀ[m0..m0+15, k0..k0+15] · Y[k0..k0+15, 0..63]  with X, Y row-major swizzled
// tiles whose ROWS are the reduction index (queries): A fragments come from ldmatrix.trans.
__device__ __forceinline__ void mma_tn(float (&acc)[8][4], const bf16* X, const bf16* Y, int m0, int k0,
                                       int lane) {
    uint32_t a0, a1, a2, a3;
    {
        const int r = k0 + (lane & 7) + ((lane >> 4) & 1) * 8;
        const int c = (m0 >> 3) + ((lane >> 3) & 1);
        ldsm_x4_t(smem_u32(X + sw_off(r, c)), a0, a1, a2, a3);
    }
    mma_x(acc, a0, a1, a2, a3, Y, k0, lane);
}

__global__ void __launch_bounds__(128, 3)
attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dctx,
                    const bf16* __restrict__ scrP, const bf16* __restrict__ scrS,
                    bf16* __restrict__ dqkv, int L, int LP, int nh, int H) {
    pdl_sync();
    extern __shared__ __align__(128) uint8_t smem_raw[];
    bf16* sQ = reinterpret_cast<bf16*>(smem_raw);  // [LP][64]
    bf16* sdO = sQ + LP * HD;                      // [LP][64]
    bf16* sP = sdO + LP * HD;                      // [LP queries][64 keys of this tile]
    bf16* sS = sP + LP * HD;

    const int bh = blockIdx.x, b = bh / nh, h = bh - b * nh;
    const int t0 = blockIdx.y * TQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = 3 * H;
    const bf16* base = qkv + (size_t)b * L * ld + h * HD;
    const bf16* dO = dctx + (size_t)b * L * H + h * HD;
    // all four tiles are indexed by QUERY row (the reduction index here): load them in 64-row
    // chunks, one cp.async group each, and start the products on chunk 0 while the rest is in flight.
    // key columns t0..t0+63 of every query row; columns >= LP (last tile) are zero-filled
    const int vch = min(8, (LP - t0) >> 3);
    const int nchunks = (LP + 63) >> 6;
    for (int c = 0; c < nchunks; ++c) {
        const int rb = c * 64, re = min(LP, rb + 64);
        load_rows(sQ, base, ld, L, rb, re, tid, 128);
        load_rows(sdO, dO, H, L, rb, re, tid, 128);
        load_rows(sP, scrP + (size_t)bh * L * LP + t0, LP, L, rb, re, tid, 128, vch);
        load_rows(sS, scrS + (size_t)bh * L * LP + t0, LP, L, rb, re, tid, 128, vch);
        load_commit();
    }
    const int r0 = t0 + warp * 16;
    const bool active = r0 < L;  // warps past the end of the sequence only take part in the barriers
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
        dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
    }
    for (int k0 = 0; k0 < LP; k0 += 16) {
        if ((k0 & 63) == 0) {
            load_wait_pending(nchunks - 1 - (k0 >> 6));
            __syncthreads();
        }
        if (!active) continue;
        mma_tn(dv, sP, sdO, warp * 16, k0, lane);  // dV += Pdᵀ·dO
        mma_tn(dk, sS, sQ, warp * 16, k0, lane);   // dK += dSᵀ·Q
    }
    if (!active) return;
    const int g = lane >> 2, t2 = (lane & 3) * 2;
    const int x0 = r0 + g, x1 = r0 + g + 8;
    bf16* dbase = dqkv + (size_t)b * L * ld + h * HD;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        if (x0 < L) {
            *reinterpret_cast<uint32_t*>(dbase + (size_t)x0 * ld + H + nt * 8 + t2) = pack_bf16(dk[nt][0], dk[nt][1]);
            *reinterpret_cast<uint32_t*>(dbase + (size_t)x0 * ld + 2 * H + nt * 8 + t2) = pack_bf16(dv[nt][0], dv[nt][1]);
        }
        if (x1 < L) {
            *reinterpret_cast<uint32_t*>(dbase + (size_t)x1 * ld + H + nt * 8 + t2) = pack_bf16(dk[nt][2], dk[nt][3]);
            *reinterpret_cast<uint32_t*>(dbase + (size_t)x1 * ld + 2 * H + nt * 8 + t2) = pack_bf16(dv[nt][2], dv[nt][3]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Backward, fused variant for joint sequences up to 176 (the C2 shape: L = 164): ONE CTA per
// (sample, head), ceil(L/16) warps. Q, K, V, dO and the whole Pd / dS matrices (bf16, three
// 64-key column tiles each) live in shared memory (222 KB), so nothing is recomputed, nothing goes
// through a global scratch buffer and every input row is read from L2 exactly once:
//   phase 1 (warp = 16 query rows): S, dPd in 32-key chunks -> Pd, dS tiles in smem ; dQ = dS·K
//   phase 2 (warp = 16 key rows)  : dV = Pdᵀ·dO ; dK = dSᵀ·Q  (ldmatrix.trans over the smem tiles)
// ---------------------------------------------------------------------------------------
// Column sums (over this warp's 16 rows) of a 16x64 accumulator tile, taken on the bf16-rounded
// values the stores write; lanes 0-3 end up with the sums of columns nt*8 + 2*lane + {0,1}.
__device__ __forceinline__ void frag_colsum(const float (&a)[8][4], float (&c)[8][2]) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        float2 lo = unpack_bf16(pack_bf16(a[nt][0], a[nt][1]));
        float2 hi = unpack_bf16(pack_bf16(a[nt][2], a[nt][3]));
        float x = lo.x + hi.x, y = lo.y + hi.y;
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            x += __shfl_xor_sync(0xffffffffu, x, o);
            y += __shfl_xor_sync(0xffffffffu, y, o);
        }
        c[nt][0] = x;
        c[nt][1] = y;
    }
}

constexpr int FUSED_MAX_LP = 176;

__global__ void __launch_bounds__(FUSED_MAX_LP * 2, 1)
attn_bwd_fused_kernel(const bf16* __restrict__ qkv, const float* __restrict__ mask,
                      const bf16* __restrict__ ctx, const bf16* __restrict__ dctx,
                      const float* __restrict__ lse, bf16* __restrict__ dqkv,
                      float* __restrict__ dbias, int L, int LP, int nh, int H, DropoutCfg drop) {
    pdl_sync();
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int TILE = LP * HD;  // elements of one [LP][64] tile
    bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
    bf16* sK = sQ + TILE;
    bf16* sV = sK + TILE;
    bf16* sdO = sV + TILE;
    bf16* sP = sdO + TILE;      // 3 tiles: keys 0-63 | 64-127 | 128-191, rows = queries
    bf16* sS = sP + 3 * TILE;   // 3 tiles, same layout
    float* sM2 = reinterpret_cast<float*>(sS + 3 * TILE);  // additive mask * log2(e), -inf past L (192 entries)
    float* sL2 = sM2 + 192;                                // lse * log2(e) per query row
    float* sD = sL2 + LP;                                  // D_i = sum_d dO[i,d] O[i,d]

    const int bh = blockIdx.x, b = bh / nh, h = bh - b * nh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x;
    const int ld = 3 * H;
    const bf16* base = qkv + (size_t)b * L * ld + h * HD;
    const bf16* dO = dctx + (size_t)b * L * H + h * HD;
    const bf16* O = ctx + (size_t)b * L * H + h * HD;

    // group 0: Q, dO and the first 64 keys of K/V; then one cp.async group per further 64-key chunk
    const int nchunks = (LP + 63) >> 6;
    load_rows(sQ, base, ld, L, 0, LP, tid, nthreads);
    load_rows(sdO, dO, H, L, 0, LP, tid, nthreads);
    for (int c = 0; c < nchunks; ++c) {
        const int rb = c * 64, re = min(LP, rb + 64);
        load_rows(sK, base + H, ld, L, rb, re, tid, nthreads);
        load_rows(sV, base + 2 * H, ld, L, rb, re, tid, nthreads);
        load_commit();
    }
    for (int j = tid; j < 192; j += nthreads) sM2[j] = (j < L) ? mask[(size_t)b * L + j] * LOG2E : -INFINITY;
    for (int i = tid >> 3; i < LP; i += nthreads >> 3) {  // D_i and lse_i: 8 threads per row
        float d = 0.f;
        if (i < L) {
            const int ch = tid & 7;
            uint4 ov = *reinterpret_cast<const uint4*>(O + (size_t)i * H + ch * 8);
            uint4 dv = *reinterpret_cast<const uint4*>(dO + (size_t)i * H + ch * 8);
            const uint32_t* op = &ov.x;
            const uint32_t* dp = &dv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float2 x = unpack_bf16(op[k]), y = unpack_bf16(dp[k]);
                d += x.x * y.x + x.y * y.y;
            }
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        if ((tid & 7) == 0) {
            sD[i] = d;
            sL2[i] = (i < L) ? lse[(size_t)bh * L + i] * LOG2E : 0.f;
        }
    }

    // ---------------- phase 1: this warp's 16 query rows against all keys ----------------
    const int r0 = warp * 16;
    const uint32_t key = drop.thresh16 ? attn_key(load_seed(drop), drop.stream) : 0u;
    const int g = lane >> 2, t2 = (lane & 3) * 2;
    const int x0 = r0 + g, x1 = r0 + g + 8;
    const bool v0 = x0 < L, v1 = x1 < L;
    uint32_t qa[4][4], da[4][4];
    float la = 0.f, lb = 0.f, Da = 0.f, Db = 0.f;
    const uint32_t pb0 = attn_block_base(bh, x0, L), pb1 = attn_block_base(bh, x1, L);
    constexpr float SC = 0.125f * LOG2E;
    float dq[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
    for (int c0 = 0; c0 < LP; c0 += 32) {
        if ((c0 & 63) == 0) {
            load_wait_pending(nchunks - 1 - (c0 >> 6));
            __syncthreads();
            if (c0 == 0) {
                load_a_frags(sQ, r0, lane, qa);
                load_a_frags(sdO, r0, lane, da);
                la = sL2[x0]; lb = sL2[x1];
                Da = sD[x0]; Db = sD[x1];
            }
        }
        // (when LP % 32 == 16 the last chunk also writes 16 all-zero key columns past LP: they stay
        //  inside the 64-wide tile and are never read by phase 2)
        bf16* tP = sP + (c0 >> 6) * TILE;
        bf16* tS = sS + (c0 >> 6) * TILE;
        float s[4][4], dp[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
        }
#pragma unroll
        for (int np = 0; np < 2; ++np) {
            if (c0 + np * 16 < LP) {
                mma_xt(s[2 * np], s[2 * np + 1], qa, sK, c0 + np * 16, lane);    // Q·Kᵀ
                mma_xt(dp[2 * np], dp[2 * np + 1], da, sV, c0 + np * 16, lane);  // dO·Vᵀ
            }
        }
#pragma unroll
        for (int np = 0; np < 2; ++np) {
            uint32_t dsp[2][2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int nt = 2 * np + e;
                const int j = c0 + nt * 8 + t2;
                const float ma = sM2[j], mb = sM2[j + 1];
                // rows / columns outside the sequence contribute exact zeros (mask = -inf past L)
                const float p0 = v0 ? ex2_approx(fmaf(s[nt][0], SC, ma) - la) : 0.f;
                const float p1 = v0 ? ex2_approx(fmaf(s[nt][1], SC, mb) - la) : 0.f;
                const float p2 = v1 ? ex2_approx(fmaf(s[nt][2], SC, ma) - lb) : 0.f;
                const float p3 = v1 ? ex2_approx(fmaf(s[nt][3], SC, mb) - lb) : 0.f;
                float k0 = 1.f, k1 = 1.f, k2 = 1.f, k3 = 1.f;
                if (drop.thresh16) {
                    const uint32_t jp = (uint32_t)j >> 1;
                    const uint32_t h0 = attn_rng(key, pb0 + jp, x0 & 1), h1 = attn_rng(key, pb1 + jp, x1 & 1);
                    k0 = ((h0 & 0xffffu) >= drop.thresh16) ? drop.scale : 0.f;
                    k1 = ((h0 >> 16) >= drop.thresh16) ? drop.scale : 0.f;
                    k2 = ((h1 & 0xffffu) >= drop.thresh16) ? drop.scale : 0.f;
                    k3 = ((h1 >> 16) >= drop.thresh16) ? drop.scale : 0.f;
                }
                dsp[e][0] = pack_bf16(p0 * (dp[nt][0] * k0 - Da) * 0.125f, p1 * (dp[nt][1] * k1 - Da) * 0.125f);
                dsp[e][1] = pack_bf16(p2 * (dp[nt][2] * k2 - Db) * 0.125f, p3 * (dp[nt][3] * k3 - Db) * 0.125f);
                const int cc = j & 63;
                const int o0 = sw_off(x0, cc >> 3) + (cc & 7), o1 = sw_off(x1, cc >> 3) + (cc & 7);
                *reinterpret_cast<uint32_t*>(tP + o0) = pack_bf16(p0 * k0, p1 * k1);
                *reinterpret_cast<uint32_t*>(tP + o1) = pack_bf16(p2 * k2, p3 * k3);
                *reinterpret_cast<uint32_t*>(tS + o0) = dsp[e][0];
                *reinterpret_cast<uint32_t*>(tS + o1) = dsp[e][1];
            }
            if (c0 + np * 16 < LP) mma_x(dq, dsp[0][0], dsp[0][1], dsp[1][0], dsp[1][1], sK, c0 + np * 16, lane);
        }
    }
    bf16* dbase = dqkv + (size_t)b * L * ld + h * HD;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        if (v0) *reinterpret_cast<uint32_t*>(dbase + (size_t)x0 * ld + nt * 8 + t2) = pack_bf16(dq[nt][0], dq[nt][1]);
        if (v1) *reinterpret_cast<uint32_t*>(dbase + (size_t)x1 * ld + nt * 8 + t2) = pack_bf16(dq[nt][2], dq[nt][3]);
    }
    float cq[8][2];
    if (dbias) frag_colsum(dq, cq);
    __syncthreads();  // every warp's rows of Pd / dS are in shared memory (and K / V are dead)
    float* sCol = reinterpret_cast<float*>(sK);  // [warps][3][64] partial bias-gradient column sums
    if (dbias && lane < 4) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            sCol[(warp * 3 + 0) * 64 + nt * 8 + t2] = cq[nt][0];
            sCol[(warp * 3 + 0) * 64 + nt * 8 + t2 + 1] = cq[nt][1];
        }
    }

    // ---------------- phase 2: this warp's 16 key rows, reduction over all queries ----------------
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
        dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
    }
    const bf16* tP2 = sP + (warp >> 2) * TILE;
    const bf16* tS2 = sS + (warp >> 2) * TILE;
    const int m0 = (warp & 3) * 16;
    for (int k0 = 0; k0 < LP; k0 += 16) {
        mma_tn(dv, tP2, sdO, m0, k0, lane);  // dV += Pdᵀ·dO
        mma_tn(dk, tS2, sQ, m0, k0, lane);   // dK += dSᵀ·Q
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        if (v0) {
            *reinterpret_cast<uint32_t*>(dbase + (size_t)x0 * ld + H + nt * 8 + t2) = pack_bf16(dk[nt][0], dk[nt][1]);
            *reinterpret_cast<uint32_t*>(dbase + (size_t)x0 * ld + 2 * H + nt * 8 + t2) = pack_bf16(dv[nt][0], dv[nt][1]);
        }
        if (v1) {
            *reinterpret_cast<uint32_t*>(dbase + (size_t)x1 * ld + H + nt * 8 + t2) = pack_bf16(dk[nt][2], dk[nt][3]);
            *reinterpret_cast<uint32_t*>(dbase + (size_t)x1 * ld + 2 * H + nt * 8 + t2) = pack_bf16(dv[nt][2], dv[nt][3]);
        }
    }
    if (dbias) {
        float ck[8][2], cv[8][2];
        frag_colsum(dk, ck);
        frag_colsum(dv, cv);
        if (lane < 4) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                sCol[(warp * 3 + 1) * 64 + nt * 8 + t2] = ck[nt][0];
                sCol[(warp * 3 + 1) * 64 + nt * 8 + t2 + 1] = ck[nt][1];
                sCol[(warp * 3 + 2) * 64 + nt * 8 + t2] = cv[nt][0];
                sCol[(warp * 3 + 2) * 64 + nt * 8 + t2 + 1] = cv[nt][1];
            }
        }
        __syncthreads();
        const int nw = nthreads >> 5;
        for (int c = tid; c < 192; c += nthreads) {  // (matrix, column) = (c / 64, c % 64)
            float acc = 0.f;
            for (int w = 0; w < nw; ++w) acc += sCol[w * 192 + c];
            atomicAdd(dbias + (c >> 6) * H + h * HD + (c & 63), acc);
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

static int launch_fwd(const void* qkv, const float* mask, void* ctx, float* lse, int B, int L, int nh,
                      int H, DropoutCfg dc, cudaStream_t stream) {
    const int LP = (L + 15) / 16 * 16;
    const size_t smem = (size_t)(2 * LP + TQ) * HD * 2 + (size_t)((L + 63) & ~63) * 4;
    if (int rc = ensure_dyn_smem((const void*)attn_fwd_kernel, smem)) return rc;
    launch_k(attn_fwd_kernel, dim3(dim3(B * nh, (L + TQ - 1) / TQ)), dim3(128), smem, stream, (const bf16*)qkv, mask, (bf16*)ctx, lse, L, LP, nh, H, dc);
    B200U_CHECK_LAUNCH("attn_fwd_kernel");
    return B200U_OK;
}

static int launch_bwd(const void* qkv, const float* mask, const void* ctx, const void* dctx,
                      const float* lse, void* dqkv, void* scratch, float* dbias, int B, int L, int nh, int H,
                      DropoutCfg dc, cudaStream_t stream) {
    const int LP = (L + 15) / 16 * 16;
    bf16* scrP = (bf16*)scratch;
    bf16* scrS = scrP + (size_t)B * nh * L * LP;
    const size_t smem1 = (size_t)(2 * LP + 2 * TQ) * HD * 2 + (size_t)(LP + 2 * TQ) * 4;
    const size_t smem2 = (size_t)4 * LP * HD * 2;
    if (int rc = ensure_dyn_smem((const void*)attn_bwd_dq_kernel, smem1)) return rc;
    if (int rc = ensure_dyn_smem((const void*)attn_bwd_dkv_kernel, smem2)) return rc;
    if (LP <= FUSED_MAX_LP) {
        // one CTA per (sample, head): Q, K, V, dO (4 tiles), Pd and dS (3 tiles each) + mask / lse / D
        const size_t smem = (size_t)10 * LP * HD * 2 + (size_t)(192 + 2 * LP) * 4;
        if (int rc = ensure_dyn_smem((const void*)attn_bwd_fused_kernel, smem)) return rc;
        launch_k(attn_bwd_fused_kernel, dim3(B * nh), dim3(LP * 2), smem, stream, (const bf16*)qkv, mask, (const bf16*)ctx, (const bf16*)dctx, lse, (bf16*)dqkv, dbias, L, LP, nh, H, dc);
        B200U_CHECK_LAUNCH("attn_bwd_fused_kernel");
        return B200U_OK;
    }
    const dim3 grid(B * nh, (L + TQ - 1) / TQ);
    launch_k(attn_bwd_dq_kernel, dim3(grid), dim3(128), smem1, stream, (const bf16*)qkv, mask, (const bf16*)ctx, (const bf16*)dctx, lse, (bf16*)dqkv, scrP, scrS, L, LP, nh, H, dc);
    B200U_CHECK_LAUNCH("attn_bwd_dq_kernel");
    launch_k(attn_bwd_dkv_kernel, dim3(grid), dim3(128), smem2, stream, (const bf16*)qkv, (const bf16*)dctx, scrP, scrS, (bf16*)dqkv, L, LP, nh, H);
    B200U_CHECK_LAUNCH("attn_bwd_dkv_kernel");
    if (dbias) return b200u_colsum_accum(dqkv, 3 * H, dbias, B * L, 3 * H, stream);
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
    if (g_attn_tc) {
        B200U_CHECK_ARG(((uintptr_t)qkv & 15) == 0 && H % 8 == 0, "attention_fwd: qkv must be 16-byte aligned");
        return attention_fwd_tc(qkv, mask, ctx, lse, B, L, num_heads, H, drop, stream);
    }
    return launch_fwd(qkv, mask, ctx, lse, B, L, num_heads, H, dc, stream);
}

extern "C" int b200u_set_attention_impl(int tcgen05) {
    g_attn_tc = tcgen05 ? 1 : 0;
    return B200U_OK;
}

extern "C" size_t b200u_attention_bwd_scratch_bytes(int B, int L, int num_heads) {
    const size_t LP = (size_t)(L + 15) / 16 * 16;
    return (size_t)2 * B * num_heads * L * LP * sizeof(bf16);
}

extern "C" int b200u_attention_bwd(const void* qkv, const float* mask, const void* ctx,
                                   const void* dctx, const float* lse, void* dqkv, void* scratch,
                                   float* dbias_qkv, int B, int L, int num_heads, int H,
                                   const b200u_dropout_t* drop, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(qkv && mask && ctx && dctx && lse && dqkv && scratch, "attention_bwd: null pointer");
    B200U_CHECK_ARG(num_heads > 0 && H == num_heads * HD, "attention_bwd: head dim must be 64 (H=%d heads=%d)", H, num_heads);
    B200U_CHECK_ARG(L > 0 && L <= 256, "attention_bwd: joint sequence length %d unsupported (1..256)", L);
    if (B == 0) return B200U_OK;
    DropoutCfg dc = make_drop(drop);
    B200U_CHECK_ARG(dc.thresh16 == 0 || dc.seed_ptr, "attention_bwd: dropout needs seed_ptr");
    if (g_attn_tc && L <= 192 && ((uintptr_t)qkv & 15) == 0 && ((uintptr_t)dctx & 15) == 0 && ((uintptr_t)ctx & 15) == 0 &&
        H % 8 == 0)
        return attention_bwd_tc(qkv, mask, ctx, dctx, lse, dqkv, dbias_qkv, B, L, num_heads, H, drop, stream);
    return launch_bwd(qkv, mask, ctx, dctx, lse, dqkv, scratch, dbias_qkv, B, L, num_heads, H, dc, stream);
}
