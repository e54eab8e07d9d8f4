// K4 (Blackwell-native) — masked-softmax self-attention forward on tcgen05 tensor cores with TMEM
// accumulators and TMA tile loads (sm_100a). Replaces model/layer.py:80-100 after the Q/K/V projections:
//   scores = Q·Kᵀ ; scores /= sqrt(64) ; scores += additive_mask ; P = softmax(scores) ; P = dropout(P)
//   ctx = P·V ; heads merged back to [B, L, H]
//
// One CTA per (sample, head, 128-query tile); the joint sequence is at most 256 long, so the whole K and V
// of the head are one TMA-loaded, 128B-swizzled operand each and a tile needs exactly two MMA groups:
//   warp 0   one elected thread: TMA loads (Q tile, K, V as 64-row boxes of the fused qkv activation),
//            S = Q·Kᵀ (M128 x N=LK x K64, 4 tcgen05.mma, accumulator in TMEM), then O = P·V
//            (M128 x N64 x K=LK) once the softmax warps have written P
//   warp 1   TMEM allocation
//   warps 2-9  softmax: THREAD = QUERY ROW x KEY HALF. A thread reads its half row of S from TMEM (tcgen05.ld):
//            the row max / sum need no shuffles (one shared-memory exchange between the two halves) and 32
//            independent scores are in flight per thread; bf16
//            probabilities go to shared memory in the K-major swizzled layout the P·V MMA reads
//            (overlaying the dead Q / K tiles), the context row comes back from TMEM and is stored merged.
// Shared memory is 73 KB and TMEM 256 columns at the C2 shape (L = 164), so two CTAs share an SM: one
// CTA's softmax overlaps the other's loads and MMAs. Only the row log-sum-exp is saved for the backward.
#include "../../include/b200u.h"
#include "common.cuh"

#include <mutex>

namespace b200u {
namespace atc {

constexpr int HD = 64;    // head dim (config/uniter-{base,large}.json: H / heads == 64)
constexpr int TM = 128;   // query rows per CTA = UMMA M
constexpr float LOG2E = 1.4426950408889634f;

// TMEM -> registers: this thread's lane (row), 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

struct FwdArgs {
    const float* mask;   // f32 [B, L] additive mask
    bf16* ctx;           // bf16 [B*L, H]
    float* lse;          // f32 [B*heads, L] or null
    int L, LK, KT, nh, H;
    DropoutCfg drop;
    long long* dbg;      // bring-up: 8 clock64 stamps per CTA
};
#define ATT_STAMP(slot)                                                                                  \
    do {                                                                                                 \
        if (a.dbg) a.dbg[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + (slot)] = clock64();        \
    } while (0)

// grid = (B * heads, ceil(L / 128)); 320 threads.
template <int TMEM_COLS>
__global__ void __launch_bounds__(320, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const FwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int KT = a.KT;                    // 64-key tiles covering LK keys
    const int regionA = max(KT * 16384, 16384 + KT * 8192);
    uint8_t* sQ = smem;                     // [128][64] bf16, K-major, 128B swizzle
    uint8_t* sK = smem + 16384;             // [KT*64][64]
    uint8_t* sP = smem;                     // KT tiles of [128][64] (overlays Q / K once S is complete)
    uint8_t* sV = smem + regionA;           // [KT*64][64]
    float* sM = reinterpret_cast<float*>(sV + KT * 8192);   // additive mask * log2(e), -inf past L (256 entries)
    float* sRed = sM + 256;                                 // [2][128] partial row max, [2][128] partial row sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 4 * TM);  // qk, v, s, p, o
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x, b = bh / a.nh, h = bh - b * a.nh;
    const int q0 = blockIdx.y * TM;
    const int L = a.L, LK = a.LK;
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    if (threadIdx.x == 0) ATT_STAMP(0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_init(&bars[3], 8);   // one arrival per softmax warp
        mbar_init(&bars[4], 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_S = *tmem_slot;
    const uint32_t tmem_O = tmem_S + (TMEM_COLS - HD);
    pdl_sync();
    if (threadIdx.x == 0) ATT_STAMP(1);

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA: Q tile + K on one barrier (S needs both), V on its own (only P·V needs it) ----
            const int row0 = b * L;
            mbar_arrive_expect_tx(&bars[0], (uint32_t)((2 + KT) * 8192));
            tma_load_2d(sQ, &tmQKV, &bars[0], h * HD, row0 + q0);
            tma_load_2d(sQ + 8192, &tmQKV, &bars[0], h * HD, row0 + q0 + 64);
            for (int i = 0; i < KT; ++i) tma_load_2d(sK + i * 8192, &tmQKV, &bars[0], a.H + h * HD, row0 + 64 * i);
            mbar_arrive_expect_tx(&bars[1], (uint32_t)(KT * 8192));
            for (int i = 0; i < KT; ++i) tma_load_2d(sV + i * 8192, &tmQKV, &bars[1], 2 * a.H + h * HD, row0 + 64 * i);
            // ---- S = Q · Kᵀ ----
            const uint32_t idS = make_idesc(TM, LK, false, false);
            const uint64_t dQ = make_smem_desc(smem_u32(sQ), 16, 1024);
            const uint64_t dK = make_smem_desc(smem_u32(sK), 16, 1024);
            mbar_wait(&bars[0], 0);
            tc_fence_after();
            ATT_STAMP(2);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, dQ + 2 * k, dK + 2 * k, idS, k > 0 ? 1u : 0u);
            umma_commit(&bars[2]);
            // ---- O = P · V (P: K-major tiles of 64 keys; V: [key][d] = MN-major B) ----
            const uint32_t idO = make_idesc(TM, HD, false, true);
            mbar_wait(&bars[3], 0);   // P is in shared memory (written through the generic proxy + fence)
            mbar_wait(&bars[1], 0);
            tc_fence_after();
            const int nk = LK >> 4;
            for (int kk = 0; kk < nk; ++kk) {
                const uint64_t dP = make_smem_desc(smem_u32(sP) + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
                const uint64_t dV = make_smem_desc(smem_u32(sV) + kk * 2048, 8192, 1024);
                umma_bf16(tmem_O, dP, dV, idO, kk > 0 ? 1u : 0u);
            }
            umma_commit(&bars[4]);
        }
        __syncwarp();
    } else if (warp >= 2) {
        // ================= softmax: thread = (query row, key half) =================
        // Two warps share a TMEM lane quadrant and split the row's keys in halves (partial max / sum are
        // combined through shared memory), so a 128-row tile keeps 8 warps busy instead of 4.
        const int q = warp & 3;                 // TMEM lane quadrant of this warp
        const int half = (warp - 2) >> 2;       // which half of the keys
        const int r = q * 32 + lane;            // row inside the tile
        const int i = q0 + r;                   // query index inside the sample
        const bool warp_live = q0 + q * 32 < L; // no valid row in this warp: only keep the barrier protocol
        for (int j = threadIdx.x - 64; j < 256; j += 256)
            sM[j] = (j < L) ? a.mask[(size_t)b * L + j] * LOG2E : -INFINITY;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const uint32_t key = a.drop.thresh16 ? attn_key(load_seed(a.drop), a.drop.stream) : 0u;
        const uint32_t pb = attn_block_base(bh, i < L ? i : 0, L);
        const uint32_t thr = a.drop.thresh16;
        constexpr float SC = 0.125f * LOG2E;    // scores / sqrt(64) (model/layer.py:86), in log2 units
        const int ksplit = (((LK >> 4) + 1) >> 1) << 4;
        const int ka = half ? ksplit : 0, kb = half ? LK : ksplit;
        mbar_wait(&bars[2], 0);
        tc_fence_after();
        if (threadIdx.x == 64) ATT_STAMP(3);
        float m = -INFINITY, l = 0.f;
        const uint32_t trow = tmem_S + ((uint32_t)(q * 32) << 16);
        if (warp_live) {
            // pass 1: row maximum of (q.k / 8 + mask) * log2(e) over this thread's keys
            for (int c0 = ka; c0 < kb; c0 += 32) {
                if (c0 + 32 <= kb) {
                    uint32_t v[32];
                    tmem_ld_32x32(trow + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 mk = *reinterpret_cast<const float4*>(sM + c0 + j);
                        m = fmaxf(m, fmaxf(fmaxf(fmaf(__uint_as_float(v[j]), SC, mk.x), fmaf(__uint_as_float(v[j + 1]), SC, mk.y)),
                                           fmaxf(fmaf(__uint_as_float(v[j + 2]), SC, mk.z), fmaf(__uint_as_float(v[j + 3]), SC, mk.w))));
                    }
                } else {
                    uint32_t v[16];
                    tmem_ld_32x16(trow + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 mk = *reinterpret_cast<const float4*>(sM + c0 + j);
                        m = fmaxf(m, fmaxf(fmaxf(fmaf(__uint_as_float(v[j]), SC, mk.x), fmaf(__uint_as_float(v[j + 1]), SC, mk.y)),
                                           fmaxf(fmaf(__uint_as_float(v[j + 2]), SC, mk.z), fmaf(__uint_as_float(v[j + 3]), SC, mk.w))));
                    }
                }
            }
            sRed[half * TM + r] = m;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 64) ATT_STAMP(4);
        if (warp_live) {
            m = fmaxf(sRed[r], sRed[TM + r]);
            // pass 2: p = exp2(score2 - m), row sum, dropout (model/layer.py:95), bf16 P into the swizzled tiles
            const int sw = r & 7;
            auto emit8 = [&](const uint32_t* v, int j0) {   // 8 consecutive keys starting at j0 (multiple of 8)
                const float4 mk0 = *reinterpret_cast<const float4*>(sM + j0);
                const float4 mk1 = *reinterpret_cast<const float4*>(sM + j0 + 4);
                float p[8];
                p[0] = ex2_approx(fmaf(__uint_as_float(v[0]), SC, mk0.x) - m);
                p[1] = ex2_approx(fmaf(__uint_as_float(v[1]), SC, mk0.y) - m);
                p[2] = ex2_approx(fmaf(__uint_as_float(v[2]), SC, mk0.z) - m);
                p[3] = ex2_approx(fmaf(__uint_as_float(v[3]), SC, mk0.w) - m);
                p[4] = ex2_approx(fmaf(__uint_as_float(v[4]), SC, mk1.x) - m);
                p[5] = ex2_approx(fmaf(__uint_as_float(v[5]), SC, mk1.y) - m);
                p[6] = ex2_approx(fmaf(__uint_as_float(v[6]), SC, mk1.z) - m);
                p[7] = ex2_approx(fmaf(__uint_as_float(v[7]), SC, mk1.w) - m);
                l += ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
                if (thr) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t hsh = attn_rng(key, pb + (uint32_t)((j0 >> 1) + e), i & 1);
                        if ((hsh & 0xffffu) < thr) p[2 * e] = 0.f;
                        if ((hsh >> 16) < thr) p[2 * e + 1] = 0.f;
                    }
                }
                uint4 o;
                o.x = pack_bf16(p[0], p[1]); o.y = pack_bf16(p[2], p[3]);
                o.z = pack_bf16(p[4], p[5]); o.w = pack_bf16(p[6], p[7]);
                *reinterpret_cast<uint4*>(sP + (j0 >> 6) * 16384 + r * 128 + ((((j0 & 63) >> 3) ^ sw) << 4)) = o;
            };
            for (int c0 = ka; c0 < kb; c0 += 32) {
                if (c0 + 32 <= kb) {
                    uint32_t v[32];
                    tmem_ld_32x32(trow + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 8) emit8(&v[j], c0 + j);
                } else {
                    uint32_t v[16];
                    tmem_ld_32x16(trow + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j += 8) emit8(&v[j], c0 + j);
                }
            }
            sRed[(2 + half) * TM + r] = l;
        }
        // P (generic-proxy writes) must be visible to the tensor core's async-proxy reads
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[3]);
        asm volatile("bar.sync 1, 256;" ::: "memory");   // both halves' partial row sums are published
        if (threadIdx.x == 64) ATT_STAMP(5);
        if (warp_live) {
            l = sRed[2 * TM + r] + sRed[3 * TM + r];
            mbar_wait(&bars[4], 0);
            tc_fence_after();
            if (threadIdx.x == 64) ATT_STAMP(6);
            // the two warps of a quadrant split the 64 context columns
            uint32_t o0[32];
            tmem_ld_32x32(tmem_O + ((uint32_t)(q * 32) << 16) + half * 32, o0);
            tmem_ld_wait();
            if (i < L) {
                if (a.lse && half == 0) a.lse[(size_t)bh * L + i] = (m + log2f(l)) * 0.69314718055994530942f;
                const float inv = (thr ? a.drop.scale : 1.0f) / l;
                bf16* out = a.ctx + ((size_t)b * L + i) * a.H + h * HD + half * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint4 o;
                    o.x = pack_bf16(__uint_as_float(o0[j]) * inv, __uint_as_float(o0[j + 1]) * inv);
                    o.y = pack_bf16(__uint_as_float(o0[j + 2]) * inv, __uint_as_float(o0[j + 3]) * inv);
                    o.z = pack_bf16(__uint_as_float(o0[j + 4]) * inv, __uint_as_float(o0[j + 5]) * inv);
                    o.w = pack_bf16(__uint_as_float(o0[j + 6]) * inv, __uint_as_float(o0[j + 7]) * inv);
                    *reinterpret_cast<uint4*>(out + j) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_S);
    }
    if (threadIdx.x == 0) ATT_STAMP(7);
}

static DropoutCfg make_drop_tc(const b200u_dropout_t* d) {
    DropoutCfg c;
    c.seed_ptr = d ? d->seed_ptr : nullptr;
    c.stream = d ? d->stream : 0;
    const float p = d ? d->p : 0.f;
    c.thresh16 = (uint32_t)(p * 65536.0f + 0.5f);
    c.scale = 1.0f / (1.0f - p);
    return c;
}


// ---------------------------------------------------------------------------------------
// Backward on tcgen05 for joint sequences up to 192 (the C2 shape: L = 164). One CTA per (sample, head);
// Q, K, V and dO arrive as TMA boxes (64 rows x 128 B, 128B swizzle) and every matrix product runs on the
// tensor cores with TMEM accumulators. The work is organised by KEY tile (128 keys) with the TRANSPOSED
// score matrix, so that a thread owns a key row and the per-query vectors (log-sum-exp, D) are broadcast
// reads:
//   Sᵀ  = K_t · Qᵀ            [128 keys x LK queries]   (A = K tile, B = Q, both K-major)
//   dPᵀ = V_t · dOᵀ           [128 keys x LK queries]
//   16 element-wise warps (thread = key row x quarter of the queries):
//        P = exp2(Sᵀ/8 + mask - lse) ; Pd = dropmask*scale*P ; dSᵀ = P*(dropmask*scale*dPᵀ - D)   (x 1/8 at read-out)
//        bf16 Pdᵀ / dSᵀ into shared memory, queries contiguous (K-major over queries)
//   dV_t = Pdᵀ · dO           A = Pdᵀ (K-major), B = dO viewed MN-major          -> complete for the tile
//   dK_t = dSᵀ · Q            A = dSᵀ (K-major), B = Q viewed MN-major           -> complete for the tile
//   dQ  += dS_t · K_t         A = the SAME dSᵀ tile viewed MN-major (M = queries), B = K_t MN-major;
//                             accumulated in TMEM over the key tiles
// TMEM: Sᵀ [0,192) | dPᵀ [192,384) | dQ [384,512); dV_t / dK_t overlay the dead Sᵀ columns. Nothing is
// recomputed, nothing goes through global scratch, no atomics except the 192 bias-gradient columns.
// ---------------------------------------------------------------------------------------
struct BwdArgs {
    const float* mask;   // f32 [B, L] additive mask
    const bf16* ctx;     // bf16 [B*L, H] forward output (for D_i = sum_d dO[i,d] O[i,d])
    const bf16* dctx;    // bf16 [B*L, H]
    const float* lse;    // f32 [B*heads, L]
    bf16* dqkv;          // bf16 [B*L, 3H]
    float* dbias;        // f32 [3H] or null: += column sums of dqkv
    int L, LK, nh, H;
    DropoutCfg drop;
    long long* dbg;
};
#define BWD_STAMP(slot)                                                           \
    do {                                                                          \
        if (a.dbg) a.dbg[(size_t)blockIdx.x * 16 + (slot)] = clock64();           \
    } while (0)

constexpr int BWD_EW_WARPS = 16;
constexpr int BWD_THREADS = 64 + 32 * BWD_EW_WARPS;
constexpr int BWD_SMEM = 5 * 24576 + 2 * 49152 + 4 * 192 * 4 + 128;

// v[0..31] per lane -> lane l returns the sum over the warp's lanes of v[l] (31 shuffles)
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
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

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmDQKV,
                   const BwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + 24576;
    uint8_t* sV = smem + 2 * 24576;
    uint8_t* sdO = smem + 3 * 24576;
    uint8_t* sO = smem + 4 * 24576;           // forward output rows (only for D_i)
    uint8_t* sST = smem + 5 * 24576;          // dSᵀ: 3 atoms of [128 keys][64 queries]
    uint8_t* sPT = sST + 49152;               // Pdᵀ: same layout (dSᵀ's 4th atom, read for padded queries, aliases it)
    float* sM2 = reinterpret_cast<float*>(sPT + 49152);  // additive mask * log2(e), -inf past L
    float* sL2 = sM2 + 192;                   // lse * log2(e) per query, +inf past L
    float* sD8 = sL2 + 192;                   // D_i per query
    float* sCol = sD8 + 192;                  // bias-gradient column sums: dq | dk | dv
    uint64_t* bars = reinterpret_cast<uint64_t*>(sCol + 192);  // ld, s[2], ew[2], o[2], rd[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x, b = bh / a.nh, h = bh - b * a.nh;
    const int L = a.L, LK = a.LK, H = a.H;
    const int ntile = (L + TM - 1) / TM;      // key tiles == query tiles
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    if (threadIdx.x == 0) BWD_STAMP(0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
        tma_prefetch_desc(&tmO);
        tma_prefetch_desc(&tmDQKV);
        mbar_init(&bars[0], 1);
        for (int k = 0; k < 2; ++k) {
            mbar_init(&bars[1 + k], 1);
            mbar_init(&bars[3 + k], BWD_EW_WARPS);
            mbar_init(&bars[5 + k], 1);
            mbar_init(&bars[7 + k], BWD_EW_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_slot;
    const uint32_t tm_S = tm, tm_dP = tm + 192, tm_dQ = tm + 384, tm_dV = tm, tm_dK = tm + 64;
    pdl_sync();

    if (warp == 0) {
        if (lane == 0) {
            const int row0 = b * L;
            mbar_arrive_expect_tx(&bars[0], 15u * 8192u);
            for (int i = 0; i < 3; ++i) {
                tma_load_2d(sO + i * 8192, &tmO, &bars[0], h * HD, row0 + 64 * i);
                tma_load_2d(sK + i * 8192, &tmQKV, &bars[0], H + h * HD, row0 + 64 * i);
                tma_load_2d(sQ + i * 8192, &tmQKV, &bars[0], h * HD, row0 + 64 * i);
                tma_load_2d(sV + i * 8192, &tmQKV, &bars[0], 2 * H + h * HD, row0 + 64 * i);
                tma_load_2d(sdO + i * 8192, &tmDO, &bars[0], h * HD, row0 + 64 * i);
            }
            const uint32_t idST = make_idesc(TM, LK, false, false);  // K-major x K-major
            const uint32_t idNV = make_idesc(TM, HD, false, true);   // K-major A, MN-major B
            const uint32_t idDQ = make_idesc(TM, HD, true, true);    // MN-major A, MN-major B
            const uint32_t uQ = smem_u32(sQ), uK = smem_u32(sK), uV = smem_u32(sV), uO = smem_u32(sdO);
            const uint32_t uST = smem_u32(sST), uPT = smem_u32(sPT);
            mbar_wait(&bars[0], 0);
            tc_fence_after();
            BWD_STAMP(1);
            const int nq = LK >> 4;   // 16-query steps
            // descriptors are built once and advanced by constants (descriptor start address is in 16-byte
            // units: +2 = 32 B = 16 K-major elements, +128 = 2048 B = 16 rows, +1024 = one 16 KB atom)
            const uint64_t dQk = make_smem_desc(uQ, 16, 1024), dOk = make_smem_desc(uO, 16, 1024);
            const uint64_t dPTk = make_smem_desc(uPT, 16, 1024), dSTk = make_smem_desc(uST, 16, 1024);
            const uint64_t dOm = make_smem_desc(uO, 8192, 1024), dQm = make_smem_desc(uQ, 8192, 1024);
            const uint64_t dSTm = make_smem_desc(uST, 16384, 1024);
            for (int kt = 0; kt < ntile; ++kt) {
                if (kt > 0) {   // dV / dK of the previous tile have been read out of the columns Sᵀ overwrites
                    mbar_wait(&bars[7 + kt - 1], 0);
                    tc_fence_after();
                }
                const uint64_t dKt = make_smem_desc(uK + kt * 16384, 16, 1024);
                const uint64_t dVt = make_smem_desc(uV + kt * 16384, 16, 1024);
                const uint64_t dKm = make_smem_desc(uK + kt * 16384, 8192, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tm_S, dKt + 2 * k, dQk + 2 * k, idST, k > 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tm_dP, dVt + 2 * k, dOk + 2 * k, idST, k > 0 ? 1u : 0u);
                umma_commit(&bars[1 + kt]);
                mbar_wait(&bars[3 + kt], 0);   // Pdᵀ / dSᵀ are in shared memory
                tc_fence_after();
                {
                    uint64_t da = dPTk, db = dOm;
                    for (int ks = 0; ks < nq; ++ks) {
                        umma_bf16(tm_dV, da, db, idNV, ks > 0 ? 1u : 0u);
                        da += ((ks & 3) == 3) ? (1024 - 6) : 2;
                        db += 128;
                    }
                    da = dSTk; db = dQm;
                    for (int ks = 0; ks < nq; ++ks) {
                        umma_bf16(tm_dK, da, db, idNV, ks > 0 ? 1u : 0u);
                        da += ((ks & 3) == 3) ? (1024 - 6) : 2;
                        db += 128;
                    }
                }
                const int nk = (min(LK, (kt + 1) * TM) - kt * TM) >> 4;   // 16-key steps of this tile
                for (int t = 0; t < ntile; ++t) {
                    uint64_t da = dSTm + (uint64_t)(2 * t) * 1024, db = dKm;
                    for (int kk = 0; kk < nk; ++kk) {
                        umma_bf16(tm_dQ + t * HD, da, db, idDQ, (kt > 0 || kk > 0) ? 1u : 0u);
                        da += 128;
                        db += 128;
                    }
                }
                umma_commit(&bars[5 + kt]);
            }
        }
        __syncwarp();
    } else if (warp >= 2) {
        const int ew = warp - 2;
        const int q = warp & 3;            // TMEM lane quadrant
        const int cq = ew >> 2;            // quarter of the columns this warp handles
        const int r = q * 32 + lane;       // row inside a 128-row tile
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int t = threadIdx.x - 64;    // 0 .. 511
        const uint32_t thr = a.drop.thresh16;
        const uint64_t seed = load_seed(a.drop);   // issued first: its latency hides under the prologue
        // ---- prologue: mask, log-sum-exp and D_i = sum_d dO[i,d] O[i,d] per query ----
        for (int j = t; j < 192; j += 32 * BWD_EW_WARPS) {
            sM2[j] = (j < L) ? a.mask[(size_t)b * L + j] * LOG2E : -INFINITY;
            sCol[j] = 0.f;
        }
        mbar_wait(&bars[0], 0);   // dO and O tiles have landed (TMA, swizzled rows of 128 B)
        if (t < 192) {
            float d = 0.f, l2 = INFINITY;
            if (t < L) {
                const uint8_t* orow = sO + t * 128;
                const uint8_t* drow = sdO + t * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 ov = *reinterpret_cast<const uint4*>(orow + ((c ^ (t & 7)) << 4));
                    const uint4 dv = *reinterpret_cast<const uint4*>(drow + ((c ^ (t & 7)) << 4));
                    const uint32_t* ow = &ov.x;
                    const uint32_t* dw = &dv.x;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 x = unpack_bf16(ow[k]), y = unpack_bf16(dw[k]);
                        d = fmaf(x.x, y.x, fmaf(x.y, y.y, d));
                    }
                }
                l2 = a.lse[(size_t)bh * L + t] * LOG2E;
            }
            sD8[t] = d;
            sL2[t] = l2;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (t == 64) BWD_STAMP(2);
        const uint32_t key = thr ? attn_key(seed, a.drop.stream) : 0u;
        const float dscale = thr ? a.drop.scale : 1.0f;
        const int L2 = (L + 1) >> 1;
        constexpr float SC = 0.125f * LOG2E;
        const int nq16 = LK >> 4;
        const int per = (nq16 + 3) >> 2;                 // 16-query chunks per column quarter
        const int c_lo = cq * per, c_hi = min(nq16, c_lo + per);
        const int sw = r & 7;
        for (int kt = 0; kt < ntile; ++kt) {
            const int j = kt * TM + r;                   // key owned by this thread
            const bool warp_live = kt * TM + q * 32 < LK;
            const float mk = (j < 192) ? sM2[j] : -INFINITY;
            const uint32_t jsh = (j & 1) ? 0u : 16u;      // odd keys use the upper 16 bits of the hash word
            const uint32_t thr_hi = thr << 16;            // (thresh16 < 65536)
            const uint32_t blk_j = (uint32_t)(bh * L2 * L2 + (j >> 1));
            mbar_wait(&bars[1 + kt], 0);
            tc_fence_after();
            if (t == 64) BWD_STAMP(3 + 4 * kt);
            if (warp_live) {
                for (int c = c_lo; c < c_hi; ++c) {
                    const int i0 = c << 4;
                    uint32_t sv[16], dv[16];
                    tmem_ld_32x16(tm_S + lane_off + i0, sv);
                    tmem_ld_32x16(tm_dP + lane_off + i0, dv);
                    tmem_ld_wait();
                    uint32_t wp[8], ws[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int i = i0 + 2 * e;
                        const float2 l2 = *reinterpret_cast<const float2*>(sL2 + i);
                        const float2 d8 = *reinterpret_cast<const float2*>(sD8 + i);
                        const float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * e]), SC, mk) - l2.x);
                        const float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * e + 1]), SC, mk) - l2.y);
                        float k0 = dscale, k1 = dscale;
                        if (thr) {
                            uint32_t hi, lo;
                            attn_rng_block(key, blk_j + (uint32_t)((i >> 1) * L2), hi, lo);
                            // this key's 16-bit sample moved to the top half, compared against thresh << 16
                            k0 = ((hi << jsh) >= thr_hi) ? dscale : 0.f;   // query i   (even row: word hi)
                            k1 = ((lo << jsh) >= thr_hi) ? dscale : 0.f;   // query i+1 (odd row: word lo)
                        }
                        // dS is left unscaled here (the 1/8 of d(scores/8) is applied once to dQ / dK at read-out)
                        const float ds0 = p0 * fmaf(__uint_as_float(dv[2 * e]), k0, -d8.x);
                        const float ds1 = p1 * fmaf(__uint_as_float(dv[2 * e + 1]), k1, -d8.y);
                        wp[e] = pack_bf16(p0 * k0, p1 * k1);
                        ws[e] = pack_bf16(ds0, ds1);
                    }
                    const int atom = i0 >> 6, ch = (i0 & 63) >> 3;
                    uint8_t* rowP = sPT + atom * 16384 + r * 128;
                    uint8_t* rowS = sST + atom * 16384 + r * 128;
                    *reinterpret_cast<uint4*>(rowP + ((ch ^ sw) << 4)) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
                    *reinterpret_cast<uint4*>(rowP + (((ch + 1) ^ sw) << 4)) = make_uint4(wp[4], wp[5], wp[6], wp[7]);
                    *reinterpret_cast<uint4*>(rowS + ((ch ^ sw) << 4)) = make_uint4(ws[0], ws[1], ws[2], ws[3]);
                    *reinterpret_cast<uint4*>(rowS + (((ch + 1) ^ sw) << 4)) = make_uint4(ws[4], ws[5], ws[6], ws[7]);
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[3 + kt]);
            if (t == 64) BWD_STAMP(4 + 4 * kt);
            // ---- dV_t / dK_t: rows = keys of this tile, this warp's 16 columns of each ----
            mbar_wait(&bars[5 + kt], 0);
            tc_fence_after();
            if (t == 64) BWD_STAMP(5 + 4 * kt);
            if (warp_live) {
                uint32_t xv[16], xk[16];
                tmem_ld_32x16(tm_dV + lane_off + cq * 16, xv);
                tmem_ld_32x16(tm_dK + lane_off + cq * 16, xk);
                tmem_ld_wait();
                float cs[32];
                uint32_t wv[8], wk[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    wk[e] = pack_bf16(__uint_as_float(xk[2 * e]) * 0.125f, __uint_as_float(xk[2 * e + 1]) * 0.125f);
                    wv[e] = pack_bf16(__uint_as_float(xv[2 * e]), __uint_as_float(xv[2 * e + 1]));
                    const float2 fk = unpack_bf16(wk[e]), fv = unpack_bf16(wv[e]);
                    const bool ok = j < L;
                    cs[2 * e] = ok ? fk.x : 0.f; cs[2 * e + 1] = ok ? fk.y : 0.f;
                    cs[16 + 2 * e] = ok ? fv.x : 0.f; cs[16 + 2 * e + 1] = ok ? fv.y : 0.f;
                }
                // bf16 rows into two swizzled [128 x 64] staging tiles (the dead Pdᵀ atoms 0 / 1); the TMA
                // store below writes them out as full 128-byte rows and clips rows past the sample's end
                uint8_t* stK = sPT + r * 128;
                uint8_t* stV = sPT + 16384 + r * 128;
                *reinterpret_cast<uint4*>(stK + (((2 * cq) ^ sw) << 4)) = make_uint4(wk[0], wk[1], wk[2], wk[3]);
                *reinterpret_cast<uint4*>(stK + (((2 * cq + 1) ^ sw) << 4)) = make_uint4(wk[4], wk[5], wk[6], wk[7]);
                *reinterpret_cast<uint4*>(stV + (((2 * cq) ^ sw) << 4)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                *reinterpret_cast<uint4*>(stV + (((2 * cq + 1) ^ sw) << 4)) = make_uint4(wv[4], wv[5], wv[6], wv[7]);
                if (a.dbias) {
                    const float tot = warp_transpose_sum32(cs, lane);   // lane < 16: dK column, else dV column
                    atomicAdd(&sCol[(lane < 16 ? 64 : 128) + cq * 16 + (lane & 15)], tot);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[7 + kt]);
            if (t == 64) BWD_STAMP(6 + 4 * kt);
            // all rows staged -> one thread stores dK_t / dV_t (2 boxes of 64 rows each) and waits until the
            // TMA engine has read the staging tiles (the next tile's element-wise phase overwrites them)
            fence_proxy_async();
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (t == 0) {
                for (int hb = 0; hb < 2; ++hb) {
                    const int row = kt * TM + hb * 64;
                    if (row < L) {
                        tma_store_3d(&tmDQKV, sPT + hb * 8192, H + h * HD, row, b);
                        tma_store_3d(&tmDQKV, sPT + 16384 + hb * 8192, 2 * H + h * HD, row, b);
                    }
                }
                bulk_commit();
                bulk_wait_read_all();
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
        }
        // ---- dQ: rows = queries (the last tile's commit covers every MMA) ----
        for (int tq = 0; tq < ntile; ++tq) {
            const int i = tq * TM + r;
            if (tq * TM + q * 32 < L) {
                uint32_t xq[16];
                tmem_ld_32x16(tm_dQ + tq * HD + lane_off + cq * 16, xq);
                tmem_ld_wait();
                float cs[32];
                uint32_t wq[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    wq[e] = pack_bf16(__uint_as_float(xq[2 * e]) * 0.125f, __uint_as_float(xq[2 * e + 1]) * 0.125f);
                    const float2 f = unpack_bf16(wq[e]);
                    cs[2 * e] = i < L ? f.x : 0.f; cs[2 * e + 1] = i < L ? f.y : 0.f;
                    cs[16 + 2 * e] = 0.f; cs[16 + 2 * e + 1] = 0.f;
                }
                uint8_t* stQ = sPT + tq * 16384 + r * 128;
                *reinterpret_cast<uint4*>(stQ + (((2 * cq) ^ sw) << 4)) = make_uint4(wq[0], wq[1], wq[2], wq[3]);
                *reinterpret_cast<uint4*>(stQ + (((2 * cq + 1) ^ sw) << 4)) = make_uint4(wq[4], wq[5], wq[6], wq[7]);
                if (a.dbias) {
                    const float tot = warp_transpose_sum32(cs, lane);
                    if (lane < 16) atomicAdd(&sCol[cq * 16 + lane], tot);
                }
            }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (t == 0) {
            for (int hb = 0; hb < 2 * ntile; ++hb)
                if (hb * 64 < L) tma_store_3d(&tmDQKV, sPT + hb * 8192, h * HD, hb * 64, b);
            bulk_commit();
            bulk_wait_read_all();
        }
        if (t == 64) BWD_STAMP(11);
        if (a.dbias && t < 192) atomicAdd(a.dbias + (t >> 6) * H + h * HD + (t & 63), sCol[t]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tm);
    }
    if (threadIdx.x == 0) BWD_STAMP(12);
}

}  // namespace atc

using namespace atc;

long long* g_attn_dbg = nullptr;   // set by b200u_gemm_debug_stamps (bring-up)

// Host launcher (called by b200u_attention_fwd when the tcgen05 path is selected).
int attention_fwd_tc(const void* qkv, const float* mask, void* ctx, float* lse, int B, int L, int nh, int H,
                     const b200u_dropout_t* drop, cudaStream_t stream) {
    FwdArgs a;
    a.mask = mask;
    a.ctx = (bf16*)ctx;
    a.lse = lse;
    a.L = L;
    a.LK = (L + 15) / 16 * 16;
    a.KT = (a.LK + 63) / 64;
    a.nh = nh;
    a.H = H;
    a.drop = make_drop_tc(drop);
    a.dbg = g_attn_dbg;
    CUtensorMap tm;
    int rc = make_tmap(&tm, qkv, B * L, 3 * H, 3 * H, 64);
    if (rc) return rc;
    const int regionA = a.KT * 16384 > 16384 + a.KT * 8192 ? a.KT * 16384 : 16384 + a.KT * 8192;
    const size_t smem = (size_t)regionA + (size_t)a.KT * 8192 + 256 * 4 + 4 * TM * 4 + 64;
    const bool wide = a.LK > 192;   // S columns + 64 context columns must fit the TMEM allocation
    auto kern = wide ? attn_fwd_tc_kernel<512> : attn_fwd_tc_kernel<256>;
    static std::mutex mu;
    static size_t set_for[2][64] = {};
    int dev = 0;
    B200U_CHECK_CUDA(cudaGetDevice(&dev));
    B200U_CHECK_ARG(dev >= 0 && dev < 64, "attention_fwd: device ordinal %d out of range", dev);
    {
        std::lock_guard<std::mutex> lock(mu);
        if (smem > set_for[wide][dev]) {
            B200U_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set_for[wide][dev] = smem;
        }
    }
    B200U_CHECK_CUDA(launch_k(kern, dim3(B * nh, (L + TM - 1) / TM), dim3(320), smem, stream, tm, a));
    B200U_CHECK_LAUNCH("attn_fwd_tc_kernel");
    return B200U_OK;
}


// Host launcher of the tcgen05 backward (L <= 192; longer sequences take the split mma.sync kernels).
int attention_bwd_tc(const void* qkv, const float* mask, const void* ctx, const void* dctx, const float* lse,
                     void* dqkv, float* dbias, int B, int L, int nh, int H, const b200u_dropout_t* drop,
                     cudaStream_t stream) {
    BwdArgs a;
    a.mask = mask;
    a.ctx = (const bf16*)ctx;
    a.dctx = (const bf16*)dctx;
    a.lse = lse;
    a.dqkv = (bf16*)dqkv;
    a.dbias = dbias;
    a.L = L;
    a.LK = (L + 15) / 16 * 16;
    a.nh = nh;
    a.H = H;
    a.drop = make_drop_tc(drop);
    a.dbg = g_attn_dbg;
    CUtensorMap tmq, tmo, tmc;
    int rc = make_tmap(&tmq, qkv, B * L, 3 * H, 3 * H, 64);
    if (rc) return rc;
    rc = make_tmap(&tmo, dctx, B * L, H, H, 64);
    if (rc) return rc;
    rc = make_tmap(&tmc, ctx, B * L, H, H, 64);
    if (rc) return rc;
    CUtensorMap tmd;
    rc = make_tmap_3d(&tmd, dqkv, B, L, 3 * H, 3 * H, 64);
    if (rc) return rc;
    static std::mutex mu;
    static bool set_for[64] = {};
    int dev = 0;
    B200U_CHECK_CUDA(cudaGetDevice(&dev));
    B200U_CHECK_ARG(dev >= 0 && dev < 64, "attention_bwd: device ordinal %d out of range", dev);
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!set_for[dev]) {
            B200U_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
            set_for[dev] = true;
        }
    }
    B200U_CHECK_CUDA(launch_k(attn_bwd_tc_kernel, dim3(B * nh), dim3(BWD_THREADS), (size_t)BWD_SMEM, stream, tmq, tmo, tmc, tmd, a));
    B200U_CHECK_LAUNCH("attn_bwd_tc_kernel");
    return B200U_OK;
}

}  // namespace b200u
