// K3 — bf16 GEMM on tcgen05 tensor cores, TMA-fed, TMEM accumulators (sm_100a).
//
// Replaces every nn.Linear / autograd matmul on the UNITER path (reference
// model/layer.py:64-66,76-78,107,112,133,140,148,153,176 and their backward) with one
// persistent warp-specialised kernel template:
//
//   warp 0      TMA producer   (cp.async.bulk.tensor → 128B-swizzled smem ring)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma, commits to mbarriers)
//   warp 2      TMEM allocator
//   warps 4-7   epilogue       (tcgen05.ld → fused bias/GELU/dropout/residual → global)
//
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of
// tile i+1. Operands may be K-major or MN-major (UMMA descriptors handle the transposed reads),
// so forward (X·Wᵀ), dgrad (dY·W) and wgrad (dYᵀ·X) all read the tensors where they lie.
#include "../../include/b200u.h"
#include "common.cuh"

#include <cudaTypedefs.h>

#include <mutex>

namespace b200u {

struct GemmArgs {
    int M, N, K;
    int splits, kb_per_split, num_kb;
    int m_tiles, n_tiles;
    void* C; int ldc;
    void* C2; int ldc2;
    const float* bias;
    const bf16* R; int ldr;
    DropoutCfg drop;
    float* colsum;                       // EPI_MUL: += column sums of the bf16 output (bias gradient), may be null
    const float* ln_gamma; const float* ln_beta;  // EPI_BIAS_DROP_RES_LN
    float* ln_mean; float* ln_rstd; float ln_eps;
    int ln_cl;                           // EPI_BIAS_DROP_RES_LN: CTAs per cluster = N / 128 column tiles of one row block
    // EPI_CE_STATS / EPI_CE_GRAD (vocabulary GEMM fused with cross entropy)
    const long long* ce_target;          // [M] target column of each row
    float* ce_partial;                   // [M][ceil(N/128)][2]: (max, sum exp(x - max)) of each row over one column tile
    float* ce_tlogit;                    // [M]: the logit at the target column
    const float* ce_lse; const float* ce_scale;  // [M]: log-sum-exp of the row, upstream gradient of its loss
    long long* dbg;  // optional per-CTA phase timestamps (8 x int64 per CTA), bring-up only
    int dbg_mode;    // bring-up: 1 = skip the MMAs (TMA-only), 2 = skip the TMA loads (MMA-only)
};

#define DBG_STAMP(slot)                                                     \
    do {                                                                    \
        if (g.dbg) g.dbg[(size_t)blockIdx.x * 8 + (slot)] = (g.dbg_mode == 3) ? (long long)globaltimer_ns() : clock64(); \
    } while (0)

// ---------------------------------------------------------------------------------------
// Fused epilogue for one output row, 32 consecutive columns starting at col0.
// ---------------------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void epilogue_row32(const float (&acc)[32], int row, int col0,
                                               const GemmArgs& g, uint64_t seed) {
    if (row >= g.M || col0 >= g.N) return;
    const int ncols = min(32, g.N - col0);
    const size_t crow = (size_t)row * g.ldc + col0;

#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = j * 8;
        if (c >= ncols) break;
        const bool full = (c + 8 <= ncols);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc[c + i];

        if (EPI == B200U_EPI_STORE || EPI == B200U_EPI_BIAS_GELU || EPI == B200U_EPI_BIAS_DROP_RES ||
            EPI == B200U_EPI_STORE_F32 || EPI == B200U_EPI_BIAS_GELU_DG) {
            if (g.bias) {
                if (full) {
                    float4 b0 = *reinterpret_cast<const float4*>(g.bias + col0 + c);
                    float4 b1 = *reinterpret_cast<const float4*>(g.bias + col0 + c + 4);
                    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                } else {
                    for (int i = 0; i < 8 && c + i < ncols; ++i) v[i] += g.bias[col0 + c + i];
                }
            }
        }
        float r[8];
        if (EPI == B200U_EPI_BIAS_DROP_RES || EPI == B200U_EPI_ADD || EPI == B200U_EPI_DGELU || EPI == B200U_EPI_MUL) {
            const bf16* rp = g.R + (size_t)row * g.ldr + col0 + c;
            if (full) {
                uint4 u = *reinterpret_cast<const uint4*>(rp);
                float2 f;
                f = unpack_bf16(u.x); r[0] = f.x; r[1] = f.y;
                f = unpack_bf16(u.y); r[2] = f.x; r[3] = f.y;
                f = unpack_bf16(u.z); r[4] = f.x; r[5] = f.y;
                f = unpack_bf16(u.w); r[6] = f.x; r[7] = f.y;
            } else {
                for (int i = 0; i < 8; ++i) r[i] = (c + i < ncols) ? __bfloat162float(rp[i]) : 0.f;
            }
        }
        if (EPI == B200U_EPI_BIAS_DROP_RES) {
            if (g.drop.thresh16) {
                // pair index over the logical [M,N] output; col0 + c is even.
                const uint32_t pbase = (uint32_t)(((size_t)row * g.N + col0 + c) >> 1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t h = rng_pair(seed, g.drop.stream, pbase + i);
                    v[2 * i] = ((h & 0xffffu) >= g.drop.thresh16) ? v[2 * i] * g.drop.scale : 0.f;
                    v[2 * i + 1] = ((h >> 16) >= g.drop.thresh16) ? v[2 * i + 1] * g.drop.scale : 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += r[i];
        } else if (EPI == B200U_EPI_ADD) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += r[i];
        } else if (EPI == B200U_EPI_DGELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= gelu_erf_grad(r[i]);
        } else if (EPI == B200U_EPI_MUL) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i] *= r[i];
                if (g.colsum && c + i < ncols) atomicAdd(g.colsum + col0 + c + i, __bfloat162float(__float2bfloat16(v[i])));
            }
        }

        if (EPI == B200U_EPI_ATOMIC_F32) {
            float* cp = reinterpret_cast<float*>(g.C) + crow + c;
            if (full) {
                red_add_v4(cp, v[0], v[1], v[2], v[3]);
                red_add_v4(cp + 4, v[4], v[5], v[6], v[7]);
            } else {
                for (int i = 0; i < 8 && c + i < ncols; ++i) atomicAdd(cp + i, v[i]);
            }
        } else if (EPI == B200U_EPI_STORE_F32) {
            float* cp = reinterpret_cast<float*>(g.C) + crow + c;
            if (full) {
                *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(cp + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
                for (int i = 0; i < 8 && c + i < ncols; ++i) cp[i] = v[i];
            }
        } else {
            bf16* cp = reinterpret_cast<bf16*>(g.C) + crow + c;
            if (full) {
                uint4 o;
                o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
                o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
                *reinterpret_cast<uint4*>(cp) = o;
            } else {
                for (int i = 0; i < 8 && c + i < ncols; ++i) cp[i] = __float2bfloat16(v[i]);
            }
            if (EPI == B200U_EPI_BIAS_GELU_DG) {
                // C = gelu'(u) and C2 = gelu(u), both of the bf16-rounded pre-activation u
                bf16* gp = reinterpret_cast<bf16*>(g.C2) + (size_t)row * g.ldc2 + col0 + c;
                for (int i = 0; i < 8 && c + i < ncols; ++i) {
                    const float u = __bfloat162float(__float2bfloat16(v[i]));
                    cp[i] = __float2bfloat16(gelu_erf_grad(u));
                    gp[i] = __float2bfloat16(gelu_erf(u));
                }
            }
            if (EPI == B200U_EPI_BIAS_GELU) {
                bf16* gp = reinterpret_cast<bf16*>(g.C2) + (size_t)row * g.ldc2 + col0 + c;
                // GELU is applied to the bf16-rounded pre-activation so that backward
                // (which only sees the stored u) differentiates exactly what forward computed.
                float w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = gelu_erf(__bfloat162float(__float2bfloat16(v[i])));
                if (full) {
                    uint4 o;
                    o.x = pack_bf16(w[0], w[1]); o.y = pack_bf16(w[2], w[3]);
                    o.z = pack_bf16(w[4], w[5]); o.w = pack_bf16(w[6], w[7]);
                    *reinterpret_cast<uint4*>(gp) = o;
                } else {
                    for (int i = 0; i < 8 && c + i < ncols; ++i) gp[i] = __float2bfloat16(w[i]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp semantics restated).
// ---------------------------------------------------------------------------------------
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;

constexpr int EPI_WARPS = 8;                      // warps 4..11
constexpr int GEMM_THREADS = (4 + EPI_WARPS) * 32;  // 384

template <int BLOCK_N, int EPI>
struct GemmCfg {
    static constexpr bool LN = EPI == B200U_EPI_BIAS_DROP_RES_LN;
    static constexpr bool HAS_R = EPI == B200U_EPI_BIAS_DROP_RES || EPI == B200U_EPI_ADD || EPI == B200U_EPI_DGELU ||
                                  EPI == B200U_EPI_MUL || LN;
    static constexpr bool DUAL = EPI == B200U_EPI_BIAS_GELU || EPI == B200U_EPI_BIAS_GELU_DG;
    static constexpr bool CE_STATS = EPI == B200U_EPI_CE_STATS;
    static constexpr bool F32_OUT = EPI == B200U_EPI_ATOMIC_F32 || EPI == B200U_EPI_STORE_F32 || CE_STATS;
    static constexpr bool HAS_BIAS = EPI == B200U_EPI_STORE || EPI == B200U_EPI_BIAS_GELU ||
                                     EPI == B200U_EPI_BIAS_DROP_RES || EPI == B200U_EPI_STORE_F32 ||
                                     EPI == B200U_EPI_BIAS_GELU_DG || LN || CE_STATS || EPI == B200U_EPI_CE_GRAD;
    static constexpr int GROUP_COLS = F32_OUT ? 32 : 64;  // one 128-byte swizzled row per output group
    static constexpr int NUM_GROUPS = BLOCK_N / GROUP_COLS;
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // epilogues with a side-input tile write their result over it in place (same swizzled slot),
    // so they need no separate staging buffer
    static constexpr int STG_PER_WARP = HAS_R ? 0 : 4096 * (DUAL ? 2 : 1);
    static constexpr int STG_BYTES = EPI_WARPS * STG_PER_WARP;
    static constexpr int R_GROUP_BYTES = BLOCK_M * 128;  // side input: one [128 x 64] bf16 box per group
    // side-input slots: the LayerNorm epilogue keeps the whole row block (one slot per group); the others
    // ping-pong over two slots (group g uses slot g & 1: the two warps of a quadrant walk even / odd groups),
    // which leaves 256-wide tiles a 4-stage operand ring instead of 3
    static constexpr int R_SLOTS = LN ? NUM_GROUPS : (NUM_GROUPS > 2 ? 2 : NUM_GROUPS);
    static constexpr int R_BYTES = HAS_R ? R_SLOTS * R_GROUP_BYTES : 0;
    static constexpr int BIAS_BYTES = 2 * BLOCK_N * 4;  // one slot per accumulator stage
    // LayerNorm epilogue: gamma / beta slices of this CTA's columns + the row-statistics mailboxes the
    // cluster's CTAs push into: [2 buffers][8 source CTAs][2 column groups][128 rows] x (sum, M2)
    static constexpr int LN_MAX_CL = BLOCK_N == 128 ? 8 : 3;   // mailbox rows: 128-wide tiles N <= 1024, 256-wide N <= 768
    static constexpr int LN_PART_BYTES = LN ? 2 * LN_MAX_CL * 2 * BLOCK_M * 8 : 0;
    static constexpr int LN_BYTES = LN ? LN_PART_BYTES + 2 * BLOCK_N * 4 : 0;
    static constexpr int FIXED = STG_BYTES + R_BYTES + BIAS_BYTES + LN_BYTES + 256 /*barriers*/;
    static constexpr int STAGES_RAW = (232448 - FIXED) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static_assert(STAGES >= 2, "not enough shared memory for a pipelined main loop");
    static constexpr int SMEM_BYTES = FIXED + STAGES * STAGE_BYTES;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages
};

// sum of the 8 bf16 values packed in o
__device__ __forceinline__ float row_sum8(const uint4& o) {
    const float2 a = unpack_bf16(o.x), b = unpack_bf16(o.y), c = unpack_bf16(o.z), d = unpack_bf16(o.w);
    return ((a.x + a.y) + (b.x + b.y)) + ((c.x + c.y) + (d.x + d.y));
}

// Epilogue math for 8 consecutive columns of one row (bf16 outputs), on element pairs with packed
// fp32x2 instructions. acc = 8 fp32 accumulators (bit patterns from tcgen05.ld); o = the 8 bf16
// results; o2 = the second output of the dual-output GELU epilogue.
template <int EPI>
__device__ __forceinline__ void epi_math8(const uint32_t* acc, const float* bias8, uint4 rraw,
                                          const GemmArgs& g, uint64_t seed, int row, int col,
                                          uint4& o, uint4& o2) {
    f32x2 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = f2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
    if (EPI == B200U_EPI_STORE || EPI == B200U_EPI_BIAS_GELU || EPI == B200U_EPI_BIAS_DROP_RES ||
        EPI == B200U_EPI_BIAS_GELU_DG || EPI == B200U_EPI_BIAS_DROP_RES_LN || EPI == B200U_EPI_CE_GRAD) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias8 + 4);
        v[0] = f2_add(v[0], f2(b0.x, b0.y));
        v[1] = f2_add(v[1], f2(b0.z, b0.w));
        v[2] = f2_add(v[2], f2(b1.x, b1.y));
        v[3] = f2_add(v[3], f2(b1.z, b1.w));
    }
    const uint32_t rr[4] = {rraw.x, rraw.y, rraw.z, rraw.w};
    uint32_t out[4], out2[4] = {0u, 0u, 0u, 0u};
    if (EPI == B200U_EPI_BIAS_GELU) {
        // GELU of the bf16-rounded pre-activation: backward sees exactly the stored u.
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            out[i] = f2_to_bf16x2(v[i]);
            out2[i] = f2_to_bf16x2(gelu_pair(f2_from_bf16x2(out[i])));
        }
    } else if (EPI == B200U_EPI_BIAS_GELU_DG) {
        // derivative AND value of GELU at the bf16-rounded pre-activation (they share erfc / exp)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f32x2 gval;
            const f32x2 gder = gelu_both_pair(f2_from_bf16x2(f2_to_bf16x2(v[i])), gval);
            out[i] = f2_to_bf16x2(gder);
            out2[i] = f2_to_bf16x2(gval);
        }
    } else if (EPI == B200U_EPI_DGELU) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            out[i] = f2_to_bf16x2(f2_mul(v[i], gelu_grad_pair(f2_from_bf16x2(rr[i]))));
    } else if (EPI == B200U_EPI_MUL) {
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = f2_to_bf16x2(f2_mul(v[i], f2_from_bf16x2(rr[i])));
    } else if (EPI == B200U_EPI_CE_GRAD) {
        // d loss / d logit = (softmax - onehot(target)) * upstream: the logits are recomputed, never stored
        const bool ok = row < g.M;
        const float lse = ok ? g.ce_lse[row] : 0.f, sc = ok ? g.ce_scale[row] : 0.f;
        const long long tg = ok ? g.ce_target[row] : -1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float x0, x1;
            f2_get(v[i], x0, x1);
            float p0 = __expf(x0 - lse), p1 = __expf(x1 - lse);
            if ((long long)(col + 2 * i) == tg) p0 -= 1.0f;
            if ((long long)(col + 2 * i + 1) == tg) p1 -= 1.0f;
            out[i] = f2_to_bf16x2(f2(p0 * sc, p1 * sc));
        }
    } else if (EPI == B200U_EPI_ADD) {
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = f2_to_bf16x2(f2_add(v[i], f2_from_bf16x2(rr[i])));
    } else if (EPI == B200U_EPI_BIAS_DROP_RES || EPI == B200U_EPI_BIAS_DROP_RES_LN) {
        if (g.drop.thresh16) {
            // 8 consecutive columns starting at a multiple of 8 = two hash quads (pairs 2q, 2q+1 share a hash)
            const uint32_t qbase = (uint32_t)(((size_t)row * g.N + col) >> 2);
            uint32_t hw[4];
            rng_quad(seed, g.drop.stream, qbase, hw[0], hw[1]);
            rng_quad(seed, g.drop.stream, qbase + 1, hw[2], hw[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t h = hw[i];
                const float m0 = ((h & 0xffffu) >= g.drop.thresh16) ? g.drop.scale : 0.f;
                const float m1 = ((h >> 16) >= g.drop.thresh16) ? g.drop.scale : 0.f;
                out[i] = f2_to_bf16x2(f2_fma(v[i], f2(m0, m1), f2_from_bf16x2(rr[i])));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) out[i] = f2_to_bf16x2(f2_add(v[i], f2_from_bf16x2(rr[i])));
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = f2_to_bf16x2(v[i]);
    }
    o = make_uint4(out[0], out[1], out[2], out[3]);
    o2 = make_uint4(out2[0], out2[1], out2[2], out2[3]);
}

// CLUSTER == 2: the two CTAs of a cluster work on vertically adjacent output tiles (same n-tile),
// each loads half of the shared B tile and TMA-multicasts it to both, halving B's L2 traffic.
template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, int CLUSTER>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
               const __grid_constant__ CUtensorMap tmR, const GemmArgs g) {
    using Cfg = GemmCfg<BLOCK_N, EPI>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int NG = Cfg::NUM_GROUPS;
    // 128B-swizzled TMA boxes need 1024-byte aligned shared memory: the dynamic window starts on a
    // 1024-byte boundary when the kernel has no static shared memory (checked below).
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = sA + STAGES * Cfg::A_BYTES;
    uint8_t* sR = sB + STAGES * Cfg::B_BYTES;
    uint8_t* sStg = sR + Cfg::R_BYTES;
    float* sBias = reinterpret_cast<float*>(sStg + Cfg::STG_BYTES);
    float* sGam = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sBias) + Cfg::BIAS_BYTES);  // LN: gamma | beta slices
    float2* sPart = reinterpret_cast<float2*>(sGam + (Cfg::LN ? 2 * BLOCK_N : 0));                // LN: statistics mailboxes
    uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + Cfg::BIAS_BYTES + Cfg::LN_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* rfull = tempty + 2;   // [4] one per side-input group
    uint64_t* rempty = rfull + 4;   // [4]
    uint64_t* lnbar = rempty + 4;   // [2] LN: all CTAs of the cluster have pushed their statistics (per mailbox buffer)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lnbar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) DBG_STAMP(0);
    if ((smem_u32(smem) & 1023u) != 0) __trap();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        if (Cfg::DUAL || Cfg::LN) tma_prefetch_desc(&tmC2);
        if (Cfg::HAS_R) tma_prefetch_desc(&tmR);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CLUSTER);  // every CTA that multicasts into this stage must be done
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], EPI_WARPS);
        }
        for (int a = 0; a < 4; ++a) {
            mbar_init(&rfull[a], 1);
            mbar_init(&rempty[a], 4);  // the four quadrant warps that own this group
        }
        if (Cfg::LN) {
            mbar_init(&lnbar[0], 1);  // armed per tile with the byte count of the cluster's statistics exchange
            mbar_init(&lnbar[1], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    // (clusters: no CTA may multicast into / arrive on a peer's barriers before the peer initialised them)
    if (CLUSTER > 1 || Cfg::LN) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the tail of
    // the preceding kernel under programmatic dependent launch; global memory is touched from here
    pdl_sync();
    if (threadIdx.x == 0) DBG_STAMP(1);

    // work units: (m-tile group of CLUSTER tiles, n-tile, k-split); a cluster walks units together.
    // LayerNorm epilogue: a unit is one 128-row block, the cluster's CTA r owns its column tile r.
    const int cta_rank = CLUSTER > 1 ? (int)cluster_ctarank() : 0;
    const int ln_rank = Cfg::LN ? (int)cluster_ctarank() : 0;
    const int ucl = Cfg::LN ? g.ln_cl : CLUSTER;  // CTAs that walk the unit list together
    const int m_groups = (g.m_tiles + CLUSTER - 1) / CLUSTER;
    const int tiles_mn = Cfg::LN ? g.m_tiles : m_groups * g.n_tiles;
    const int total_tiles = tiles_mn * g.splits;
    const int unit0 = blockIdx.x / ucl;
    const int unit_stride = gridDim.x / ucl;
    constexpr uint16_t MC_MASK = (uint16_t)((1u << CLUSTER) - 1);
    // origin of the output tile of unit `rem` (index inside one k-split)
    auto tile_origin = [&](int rem, int& m0, int& n0) {
        if (Cfg::LN) {
            m0 = rem * BLOCK_M;
            n0 = ln_rank * BLOCK_N;
        } else {
            m0 = ((rem % m_groups) * CLUSTER + cta_rank) * BLOCK_M;
            n0 = (rem / m_groups) * BLOCK_N;
        }
    };

    if (warp == 0) {
        // ================= TMA producer =================
        // The whole warp walks the loop with warp-uniform state and one elected lane issues the
        // barrier arrivals and TMA requests: with uniform operands the requests compile to plain
        // uniform-datapath UTMALDGs (no per-request R2UR + elect waterfall), which matters because at
        // 128-wide tiles this single instruction stream must sustain one k-block every ~250 cycles.
        const bool leader = elect_one();
        const uint32_t sA_u = __shfl_sync(0xffffffffu, smem_u32(sA), 0);
        const uint32_t sB_u = __shfl_sync(0xffffffffu, smem_u32(sB), 0);
        const uint32_t sR_u = __shfl_sync(0xffffffffu, smem_u32(sR), 0);
        const uint32_t full_u = __shfl_sync(0xffffffffu, smem_u32(full), 0);
        const uint32_t rfull_u = __shfl_sync(0xffffffffu, smem_u32(rfull), 0);
        int s = 0;
        uint32_t ph = 0;
        // Side-input cursor: the next (tile, group) whose R box has not been requested. Boxes are
        // requested opportunistically (try_wait on the group's empty barrier) from inside the
        // operand loop and from its wait loops, never by blocking it: the epilogue that frees a
        // group may itself be waiting, through the MMA warp, on operands only this warp loads.
        int rt = unit0, rg = 0;
        uint32_t rpar = 0;  // bit g: parity of the next use of group g's barriers
        int t_cur = unit0;
        int rm0 = 0, rn0 = 0;  // tile origin of the cursor (divisions only when the cursor moves)
        auto r_origin = [&]() { tile_origin(rt % tiles_mn, rm0, rn0); };
        if (Cfg::HAS_R) r_origin();
        auto r_pump = [&](bool block) {
            if (!Cfg::HAS_R) return;
            while (rt <= t_cur && rt < total_tiles) {
                if (rn0 + rg * 64 < g.N) {
                    const int rs = rg % Cfg::R_SLOTS;
                    const uint32_t par = ((rpar >> rs) & 1u) ^ 1u;
                    if (block) {
                        mbar_wait(&rempty[rs], par);
                    } else {
                        const int ok = __shfl_sync(0xffffffffu, (int)mbar_test_wait(&rempty[rs], par), 0);
                        if (!ok) return;
                    }
                    if (leader) {
                        mbar_arrive_expect_tx_u(rfull_u + rs * 8, Cfg::R_GROUP_BYTES);
                        tma_load_2d_u(sR_u + rs * Cfg::R_GROUP_BYTES, &tmR, rfull_u + rs * 8, rn0 + 64 * rg, rm0);
                    }
                    __syncwarp();
                    rpar ^= 1u << rs;
                }
                if (++rg == NG) {
                    rg = 0;
                    rt += unit_stride;
                    r_origin();
                }
            }
        };
        for (int t = unit0; t < total_tiles; t += unit_stride) {
            t_cur = t;
            const int split = t / tiles_mn;
            const int rem = t - split * tiles_mn;
            int m0, n0;
            tile_origin(rem, m0, n0);
            const int kb0 = split * g.kb_per_split;
            const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
            for (int kb = kb0; kb < kb1; ++kb) {
                if (Cfg::HAS_R) {
                    while (!mbar_test_wait(&empty[s], ph ^ 1)) r_pump(false);
                } else {
                    mbar_wait(&empty[s], ph ^ 1);
                }
                if (leader) {
                    const uint32_t fb = full_u + s * 8;
                    if (g.dbg_mode == 2) mbar_arrive(&full[s]);
                    else mbar_arrive_expect_tx_u(fb, Cfg::STAGE_BYTES);
                    const uint32_t a_dst = sA_u + s * Cfg::A_BYTES;
                    const uint32_t b_dst = sB_u + s * Cfg::B_BYTES;
                    if (g.dbg_mode == 2) {
                        // bring-up: no operand traffic, the MMAs run on whatever is in smem
                    } else if (!A_MN) {
                        tma_load_2d_u(a_dst, &tmA, fb, kb * BLOCK_K, m0);
                    } else {
#pragma unroll
                        for (int i = 0; i < BLOCK_M / 64; ++i)
                            tma_load_2d_u(a_dst + i * (BLOCK_K * 128), &tmA, fb, m0 + 64 * i, kb * BLOCK_K);
                    }
                    if (g.dbg_mode == 2) {
                    } else if (CLUSTER == 1) {
                        if (!B_MN) {
                            tma_load_2d_u(b_dst, &tmB, fb, kb * BLOCK_K, n0);
                        } else {
#pragma unroll
                            for (int i = 0; i < BLOCK_N / 64; ++i)
                                tma_load_2d_u(b_dst + i * (BLOCK_K * 128), &tmB, fb, n0 + 64 * i, kb * BLOCK_K);
                        }
                    } else {
                        // this CTA fetches its 1/CLUSTER slice of the B tile and multicasts it
                        uint8_t* b_ptr = sB + s * Cfg::B_BYTES;
                        if (!B_MN) {
                            constexpr int ROWS = BLOCK_N / CLUSTER;
                            tma_load_2d_mc(b_ptr + cta_rank * (ROWS * 128), &tmB, &full[s], kb * BLOCK_K,
                                           n0 + cta_rank * ROWS, MC_MASK);
                        } else {
                            constexpr int BOXES = BLOCK_N / 64 / CLUSTER;
#pragma unroll
                            for (int j = 0; j < BOXES; ++j) {
                                const int i = cta_rank * BOXES + j;
                                tma_load_2d_mc(b_ptr + i * (BLOCK_K * 128), &tmB, &full[s], n0 + 64 * i,
                                               kb * BLOCK_K, MC_MASK);
                            }
                        }
                    }
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
                if (Cfg::HAS_R) r_pump(false);
            }
        }
        r_pump(true);  // whatever side input is still outstanding (nothing else left to load)
        if (lane == 0) DBG_STAMP(2);
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // Everything the issuing lane touches is kept warp-uniform (tile/stage counters, smem and
        // TMEM addresses broadcast with shfl) and the two smem descriptors are built once per stage
        // and advanced by adding a constant, so the SASS between consecutive UTCHMMAs is a couple of
        // uniform-datapath adds instead of a per-MMA descriptor rebuild + R2UR waterfall.
        constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, A_MN, B_MN);
        constexpr uint64_t A_TMPL = desc_template(A_MN ? BLOCK_K * 128 : 16, 1024);
        constexpr uint64_t B_TMPL = desc_template(B_MN ? BLOCK_K * 128 : 16, 1024);
        constexpr uint64_t A_STEP = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
        constexpr uint64_t B_STEP = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t a_base = __shfl_sync(0xffffffffu, smem_u32(sA), 0);
        const uint32_t b_base = __shfl_sync(0xffffffffu, smem_u32(sB), 0);
        const bool leader = elect_one();
        int s = 0;
        uint32_t ph = 0;
        int as = 0;
        uint32_t aph = 0;
        for (int t = unit0; t < total_tiles; t += unit_stride) {
            const int split = t / tiles_mn;
            const int kb0 = split * g.kb_per_split;
            const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
            mbar_wait(&tempty[as], aph ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_u + as * BLOCK_N;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (lane == 0 && kb == kb0 && t == unit0) DBG_STAMP(3);
                if (leader) {
                    if (g.dbg_mode != 1) {
                        const uint64_t ad = A_TMPL | (uint64_t)(((a_base + s * Cfg::A_BYTES) >> 4) & 0x3FFF);
                        const uint64_t bd = B_TMPL | (uint64_t)(((b_base + s * Cfg::B_BYTES) >> 4) & 0x3FFF);
                        umma_bf16(tmem_d, ad, bd, idesc, kb > kb0 ? 1u : 0u);
#pragma unroll
                        for (int k = 1; k < BLOCK_K / UMMA_K; ++k)
                            umma_bf16(tmem_d, ad + k * A_STEP, bd + k * B_STEP, idesc, 1u);
                    }
                    if (g.dbg_mode == 1) {
                        mbar_arrive(&empty[s]);  // TMA-only bring-up mode (CLUSTER == 1 only)
                        if (kb == kb1 - 1) mbar_arrive(&tfull[as]);
                    } else {
                        if (CLUSTER == 1) umma_commit(&empty[s]);
                        else umma_commit_mc(&empty[s], MC_MASK);
                        if (kb == kb1 - 1) umma_commit(&tfull[as]);
                    }
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            as ^= 1;
            if (as == 0) aph ^= 1;
        }
        if (lane == 0) DBG_STAMP(4);
    } else if (warp >= 4) {
        // ================= epilogue: TMEM -> regs -> fused math -> swizzled smem -> TMA store =========
        // TMEM is read in 32-column halves that ping-pong between two register sets, so the next
        // tcgen05.ld is always in flight while the previous half goes through the epilogue math.
        const int ew = warp - 4;
        const int q = warp & 3;    // TMEM lane quadrant this warp may access
        const int half = ew >> 2;  // the two warps of a quadrant alternate over column groups
        uint8_t* stg = sStg + ew * Cfg::STG_PER_WARP;
        const uint64_t seed = (EPI == B200U_EPI_BIAS_DROP_RES || Cfg::LN) ? load_seed(g.drop) : 0ull;
        const int rr = q * 32 + lane;  // row inside the tile
        int ln_tiles = 0;              // LN: tiles this CTA has finished (mailbox buffer / barrier parity)
        float ln_s1 = 0.f;             // LN: sum of this row's bf16 values over this warp's column groups
        const int sw = lane & 7;       // 128B-swizzle phase of this row
        int as = 0;
        uint32_t aph = 0;
        uint32_t rpar = 0;  // bit g: parity of the next fill of side-input group g
        // bias slice of the NEXT tile, fetched into registers by epilogue warp 0 while the current
        // tile is being processed, published through one smem slot per accumulator stage
        float bpre[BLOCK_N / 32];
        auto bias_fetch = [&](int tile) {
            int bm0, bn0;
            tile_origin(tile % tiles_mn, bm0, bn0);
#pragma unroll
            for (int i = 0; i < BLOCK_N / 32; ++i) {
                const int c = bn0 + lane + 32 * i;
                bpre[i] = (g.bias && c < g.N) ? g.bias[c] : 0.f;
            }
        };
        if (Cfg::HAS_BIAS && ew == 0 && unit0 < total_tiles) bias_fetch(unit0);
        if (Cfg::LN && ew == 1) {
            // this CTA's column tile never changes: gamma / beta slices are staged once (published by the
            // first tile's epilogue barrier below)
#pragma unroll
            for (int i = 0; i < BLOCK_N / 32; ++i) {
                const int c = ln_rank * BLOCK_N + lane + 32 * i;
                sGam[lane + 32 * i] = c < g.N ? g.ln_gamma[c] : 0.f;
                sGam[BLOCK_N + lane + 32 * i] = c < g.N ? g.ln_beta[c] : 0.f;
            }
        }
        for (int t = unit0; t < total_tiles; t += unit_stride) {
            const int split = t / tiles_mn;
            const int rem = t - split * tiles_mn;
            int m0, n0;
            tile_origin(rem, m0, n0);
            const float* bias_s = sBias + as * BLOCK_N;
            if (Cfg::HAS_BIAS) {
                if (ew == 0) {
#pragma unroll
                    for (int i = 0; i < BLOCK_N / 32; ++i) sBias[as * BLOCK_N + lane + 32 * i] = bpre[i];
                }
                // slot `as` was last read two tiles ago; every warp passed the previous tile's barrier since
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
                if (ew == 0 && t + unit_stride < total_tiles) bias_fetch(t + unit_stride);
            }
            mbar_wait(&tfull[as], aph);
            tc_fence_after();
            if (threadIdx.x == 128 && t == unit0) DBG_STAMP(5);
            const int row = m0 + rr;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N;
            bool released = false;  // this warp's arrival on tempty[as]
            auto release_acc = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[as]);
                released = true;
            };
            int gi = half;
            bool have = gi < NG && n0 + gi * Cfg::GROUP_COLS < g.N;
            uint32_t ra[32], rb[32];
            if (have) tmem_ld_32x32(taddr + gi * Cfg::GROUP_COLS, ra);
            if (Cfg::F32_OUT) {
                // cross-entropy statistics of this thread's row over this warp's column groups (online softmax)
                float ce_m = -INFINITY, ce_s = 0.f, ce_t = 0.f;
                bool ce_hit = false;
                long long ce_tgt = -1;
                if (Cfg::CE_STATS && row < g.M) ce_tgt = g.ce_target[row];
                // fp32 output (wgrad reduce-add / fp32 store): 32-column groups, one 4 KB staging tile
                auto stage_out = [&](const uint32_t (&r)[32], int col0) {
                    if constexpr (Cfg::CE_STATS) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int c = n0 + col0 + j;
                            const float x = __uint_as_float(r[j]) + bias_s[col0 + j];
                            if (c < g.N) {
                                if (x > ce_m) {
                                    ce_s = ce_s * __expf(ce_m - x) + 1.0f;
                                    ce_m = x;
                                } else {
                                    ce_s += __expf(x - ce_m);
                                }
                                if ((long long)c == ce_tgt) { ce_t = x; ce_hit = true; }
                            }
                        }
                        return;
                    }
                    if (lane == 0) bulk_wait_read_all();  // staging buffer free again?
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o;
                        o.x = __uint_as_float(r[4 * j + 0]); o.y = __uint_as_float(r[4 * j + 1]);
                        o.z = __uint_as_float(r[4 * j + 2]); o.w = __uint_as_float(r[4 * j + 3]);
                        if (EPI == B200U_EPI_STORE_F32) {
                            const float4 bb = *reinterpret_cast<const float4*>(bias_s + col0 + 4 * j);
                            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                        }
                        *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ sw) << 4)) = o;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (EPI == B200U_EPI_ATOMIC_F32) tma_reduce_add_2d(&tmC, stg, n0 + col0, m0 + q * 32);
                        else tma_store_2d(&tmC, stg, n0 + col0, m0 + q * 32);
                        bulk_commit();
                    }
                };
#pragma unroll 1
                while (have) {
                    int gn = gi + 2;
                    bool more = gn < NG && n0 + gn * 32 < g.N;
                    tmem_ld_wait();  // ra
                    if (more) tmem_ld_32x32(taddr + gn * 32, rb); else release_acc();
                    stage_out(ra, gi * 32);
                    if (!more) break;
                    gi = gn;
                    gn = gi + 2;
                    more = gn < NG && n0 + gn * 32 < g.N;
                    tmem_ld_wait();  // rb
                    if (more) tmem_ld_32x32(taddr + gn * 32, ra); else release_acc();
                    stage_out(rb, gi * 32);
                    gi = gn;
                    have = more;
                }
                if constexpr (Cfg::CE_STATS) {
                    // the two warps of a lane quadrant walked the even / odd column groups: merge them, then one
                    // (max, sum) pair per (row, column tile) and the target logit leave for global memory
                    float* xch = reinterpret_cast<float*>(stg);
                    if (half == 1) {
                        *reinterpret_cast<float4*>(xch + lane * 4) = make_float4(ce_m, ce_s, ce_t, ce_hit ? 1.f : 0.f);
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
                    if (half == 0) {
                        const float4 o = *reinterpret_cast<const float4*>(
                            reinterpret_cast<const float*>(sStg + (ew + 4) * Cfg::STG_PER_WARP) + lane * 4);
                        const float mm = fmaxf(ce_m, o.x);
                        float ss = 0.f;
                        if (ce_m > -INFINITY) ss += ce_s * __expf(ce_m - mm);
                        if (o.x > -INFINITY) ss += o.y * __expf(o.x - mm);
                        if (row < g.M) {
                            reinterpret_cast<float2*>(g.ce_partial)[(size_t)row * g.n_tiles + n0 / BLOCK_N] = make_float2(mm, ss);
                            if (ce_hit) g.ce_tlogit[row] = ce_t;
                            else if (o.w != 0.f) g.ce_tlogit[row] = o.z;
                        }
                    }
                }
            } else {
                int prev = -1;  // side-input group whose TMA store may still be reading its slot
#pragma unroll 1
                while (have) {
                    const int col0 = gi * 64;
                    const int gn = gi + 2;
                    const bool more = gn < NG && n0 + gn * 64 < g.N;
                    const int rs = Cfg::HAS_R ? gi % Cfg::R_SLOTS : 0;   // side-input slot of this group
                    uint8_t* slot = Cfg::HAS_R ? sR + rs * Cfg::R_GROUP_BYTES : stg;
                    uint8_t* dst = Cfg::HAS_R ? slot + q * (32 * 128) : stg;
                    tmem_ld_wait();                              // ra = columns col0 .. col0+31
                    tmem_ld_32x32(taddr + col0 + 32, rb);        // lands during the math on ra
                    // staging reuse: the previous group's TMA store must have finished reading smem
                    // (side-input epilogues write in place, a different slot per group: no wait here)
                    if (!Cfg::HAS_R && lane == 0) bulk_wait_read_all();
                    if (Cfg::HAS_R) {
                        mbar_wait(&rfull[rs], (rpar >> rs) & 1u);
                        rpar ^= 1u << rs;
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 rraw = make_uint4(0, 0, 0, 0);
                        if (Cfg::HAS_R)
                            rraw = *reinterpret_cast<const uint4*>(slot + rr * 128 + ((j ^ sw) << 4));
                        uint4 o, o2;
                        epi_math8<EPI>(&ra[8 * j], bias_s + col0 + 8 * j, rraw, g, seed, row, n0 + col0 + 8 * j, o, o2);
                        *reinterpret_cast<uint4*>(dst + lane * 128 + ((j ^ sw) << 4)) = o;
                        if (Cfg::DUAL)
                            *reinterpret_cast<uint4*>(stg + 4096 + lane * 128 + ((j ^ sw) << 4)) = o2;
                        if (Cfg::LN) ln_s1 += row_sum8(o);
                    }
                    tmem_ld_wait();                              // rb = columns col0+32 .. col0+63
                    if (more) tmem_ld_32x32(taddr + gn * 64, ra); else release_acc();
#pragma unroll
                    for (int j = 4; j < 8; ++j) {
                        uint4 rraw = make_uint4(0, 0, 0, 0);
                        if (Cfg::HAS_R)
                            rraw = *reinterpret_cast<const uint4*>(slot + rr * 128 + ((j ^ sw) << 4));
                        uint4 o, o2;
                        epi_math8<EPI>(&rb[8 * (j - 4)], bias_s + col0 + 8 * j, rraw, g, seed, row, n0 + col0 + 8 * j, o, o2);
                        *reinterpret_cast<uint4*>(dst + lane * 128 + ((j ^ sw) << 4)) = o;
                        if (Cfg::DUAL)
                            *reinterpret_cast<uint4*>(stg + 4096 + lane * 128 + ((j ^ sw) << 4)) = o2;
                        if (Cfg::LN) ln_s1 += row_sum8(o);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmC, dst, n0 + col0, m0 + q * 32);
                        if (Cfg::DUAL) tma_store_2d(&tmC2, stg + 4096, n0 + col0, m0 + q * 32);
                        bulk_commit();
                        if (Cfg::HAS_R && !Cfg::LN && Cfg::R_SLOTS == NG && prev >= 0) {
                            // all but the store just committed have been read: the previous group's
                            // side-input slot can be refilled for the next tile
                            bulk_wait_read_but_one();
                            mbar_arrive(&rempty[prev]);
                        }
                    }
                    if (EPI == B200U_EPI_MUL && g.colsum) {
                        // bias gradient of the dense whose output gradient this tile is: column sums of the
                        // bf16 results over this warp's 32 rows, read back from the swizzled tile (lane c owns
                        // columns 2c, 2c+1: the 32 lanes cover one 128-byte row, conflict-free), one fp32
                        // atomic per column and warp. Rows past M are exact zeros (operands zero-filled).
                        float c0 = 0.f, c1 = 0.f;
                        const int chunk = lane >> 2, wq = (lane & 3) * 4;
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(dst + r * 128 + ((chunk ^ (r & 7)) << 4) + wq));
                            c0 += f.x;
                            c1 += f.y;
                        }
                        const int col = n0 + col0 + 2 * lane;
                        if (col < g.N) atomicAdd(g.colsum + col, c0);
                        if (col + 1 < g.N) atomicAdd(g.colsum + col + 1, c1);
                    }
                    if (Cfg::HAS_R && !Cfg::LN && Cfg::R_SLOTS < NG) {
                        // ping-pong slots: this warp's next group reuses the slot, so it is handed back as soon
                        // as the store has read it (and, EPI_MUL, every lane's column-sum reads are done)
                        __syncwarp();
                        if (lane == 0) {
                            bulk_wait_read_all();
                            mbar_arrive(&rempty[rs]);
                        }
                    }
                    prev = gi;
                    gi = gn;
                    have = more;
                }
                if (Cfg::LN) {
                    // ---- fused LayerNorm over the full row (model/layer.py:111-115,152-156) ----
                    // This warp's NG/2 column groups (half, half + 2, ...) of the row are in shared memory as the
                    // bf16 values the stores above write (= what the backward reads): in place in the side-input
                    // slots, one row per lane.
                    constexpr int MYG = NG / 2;                       // groups per warp
                    constexpr float NLOC = 64.0f * MYG;
                    // (1) local statistics: sum (accumulated above), then M2 about the local mean
                    const float mloc = ln_s1 * (1.0f / NLOC);
                    float m2 = 0.f;
#pragma unroll
                    for (int gg = 0; gg < MYG; ++gg) {
                        const uint8_t* src = sR + (half + 2 * gg) * Cfg::R_GROUP_BYTES + rr * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint4 v4 = *reinterpret_cast<const uint4*>(src + ((j ^ sw) << 4));
                            const uint32_t* w4 = &v4.x;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 f = unpack_bf16(w4[k]);
                                m2 = fmaf(f.x - mloc, f.x - mloc, m2);
                                m2 = fmaf(f.y - mloc, f.y - mloc, m2);
                            }
                        }
                    }
                    // (2) push (sum, M2) into the mailbox [buffer][this CTA][half][row] of EVERY CTA of the cluster
                    // with st.async: each store completes 8 transaction bytes on the destination's barrier, which
                    // one local thread armed with the byte count of the whole exchange
                    const int buf = ln_tiles & 1;
                    if (g.ln_cl == 1) {
                        // one column tile covers the row: only this CTA's two halves meet, through plain shared
                        // memory and the epilogue warps' named barrier (no distributed shared memory involved)
                        sPart[((buf * Cfg::LN_MAX_CL) * 2 + half) * BLOCK_M + rr] = make_float2(ln_s1, m2);
                        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
                    } else {
                        if (ew == 0 && lane == 0) mbar_arrive_expect_tx(&lnbar[buf], (uint32_t)(g.ln_cl * 2 * BLOCK_M * 8));
                        const uint32_t mbox = smem_u32(sPart + ((buf * Cfg::LN_MAX_CL + ln_rank) * 2 + half) * BLOCK_M + rr);
                        const uint32_t lb = smem_u32(&lnbar[buf]);
                        for (int r = 0; r < g.ln_cl; ++r)
                            dsmem_st_async_f32x2(dsmem_addr(mbox, (uint32_t)r), ln_s1, m2, dsmem_addr(lb, (uint32_t)r));
                        // (3) all 2 * cluster-size partials of the row have landed here: Chan's combination
                        mbar_wait_cluster(&lnbar[buf], (uint32_t)(ln_tiles >> 1) & 1u);
                    }
                    float tot = 0.f;
                    const float2* pp = sPart + (size_t)buf * Cfg::LN_MAX_CL * 2 * BLOCK_M + rr;
                    for (int r = 0; r < 2 * g.ln_cl; ++r) tot += pp[r * BLOCK_M].x;
                    const float mean = tot / (float)g.N;
                    float M2 = 0.f;
                    for (int r = 0; r < 2 * g.ln_cl; ++r) {
                        const float2 pr = pp[r * BLOCK_M];
                        const float dm = pr.x * (1.0f / NLOC) - mean;
                        M2 += pr.y + NLOC * dm * dm;
                    }
                    const float rstd = rsqrtf(M2 / (float)g.N + g.ln_eps);
                    if (ln_rank == 0 && half == 0 && row < g.M) {
                        if (g.ln_mean) g.ln_mean[row] = mean;
                        if (g.ln_rstd) g.ln_rstd[row] = rstd;
                    }
                    // (4) normalise in place (once the stores of the pre-LN values have finished reading the
                    // slots) and store the LayerNorm output through tmC2
                    if (lane == 0) bulk_wait_read_all();
                    __syncwarp();
                    const float nmr = -mean * rstd;
#pragma unroll
                    for (int gg = 0; gg < MYG; ++gg) {
                        const int col0 = (half + 2 * gg) * 64;
                        uint8_t* dst = sR + (half + 2 * gg) * Cfg::R_GROUP_BYTES + q * (32 * 128);
                        const float* gam = sGam + col0;
                        const float* bet = sGam + BLOCK_N + col0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            uint4* cell = reinterpret_cast<uint4*>(dst + lane * 128 + ((j ^ sw) << 4));
                            const uint4 v4 = *cell;
                            const float4 ga0 = *reinterpret_cast<const float4*>(gam + 8 * j);
                            const float4 ga1 = *reinterpret_cast<const float4*>(gam + 8 * j + 4);
                            const float4 be0 = *reinterpret_cast<const float4*>(bet + 8 * j);
                            const float4 be1 = *reinterpret_cast<const float4*>(bet + 8 * j + 4);
                            const float2 f0 = unpack_bf16(v4.x), f1 = unpack_bf16(v4.y);
                            const float2 f2_ = unpack_bf16(v4.z), f3 = unpack_bf16(v4.w);
                            uint4 ov;
                            ov.x = pack_bf16(fmaf(fmaf(f0.x, rstd, nmr), ga0.x, be0.x), fmaf(fmaf(f0.y, rstd, nmr), ga0.y, be0.y));
                            ov.y = pack_bf16(fmaf(fmaf(f1.x, rstd, nmr), ga0.z, be0.z), fmaf(fmaf(f1.y, rstd, nmr), ga0.w, be0.w));
                            ov.z = pack_bf16(fmaf(fmaf(f2_.x, rstd, nmr), ga1.x, be1.x), fmaf(fmaf(f2_.y, rstd, nmr), ga1.y, be1.y));
                            ov.w = pack_bf16(fmaf(fmaf(f3.x, rstd, nmr), ga1.z, be1.z), fmaf(fmaf(f3.y, rstd, nmr), ga1.w, be1.w));
                            *cell = ov;
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int gg = 0; gg < MYG; ++gg)
                            tma_store_2d(&tmC2, sR + (half + 2 * gg) * Cfg::R_GROUP_BYTES + q * (32 * 128),
                                         n0 + (half + 2 * gg) * 64, m0 + q * 32);
                        bulk_commit();
                        bulk_wait_read_all();   // the slots may be refilled with the next tile's side input
#pragma unroll
                        for (int gg = 0; gg < MYG; ++gg) mbar_arrive(&rempty[half + 2 * gg]);
                    }
                    ++ln_tiles;
                    ln_s1 = 0.f;
                } else if (Cfg::HAS_R && Cfg::R_SLOTS == NG && prev >= 0) {
                    __syncwarp();  // (EPI_MUL: every lane's column-sum reads of the slot are done)
                    if (lane == 0) {
                        bulk_wait_read_all();  // the last store has finished reading its side-input slot
                        mbar_arrive(&rempty[prev]);
                    }
                }
            }
            if (!released) release_acc();  // warps that had no column group in this tile
            as ^= 1;
            if (as == 0) aph ^= 1;
        }
        // the TMA stores only have to be done READING shared memory before the CTA exits; their
        // global writes are complete (and visible to dependent grids) when the grid completes
        if (lane == 0) bulk_wait_read_all();
        if (threadIdx.x == 128) DBG_STAMP(6);
    }

    tc_fence_before();
    // no CTA may exit while its peer can still multicast into its smem or arrive on its barriers
    // (LayerNorm clusters: every push destined to this CTA has been waited for by its epilogue, peers
    //  never read remote memory, so a CTA may leave as soon as it is done)
    if (CLUSTER > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
    if (threadIdx.x == 0) DBG_STAMP(7);
}

// ---------------------------------------------------------------------------------------
// SIMT debug kernel: same contract and the same epilogue, no tensor cores. Selected only by
// b200u_gemm_t.impl = 1 (bring-up / differential testing of the tcgen05 path on the GPU).
// ---------------------------------------------------------------------------------------
template <int EPI>
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, int lda, int a_mn,
                                 const bf16* __restrict__ B, int ldb, int b_mn, const GemmArgs g) {
    pdl_sync();
    const int row = blockIdx.y * 128 + threadIdx.x;
    const int col0 = blockIdx.x * 32;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    if (row < g.M) {
        for (int k = 0; k < g.K; ++k) {
            const float a = __bfloat162float(a_mn ? A[(size_t)k * lda + row] : A[(size_t)row * lda + k]);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int n = col0 + i;
                if (n < g.N) {
                    const float b =
                        __bfloat162float(b_mn ? B[(size_t)k * ldb + n] : B[(size_t)n * ldb + k]);
                    acc[i] = fmaf(a, b, acc[i]);
                }
            }
        }
    }
    const uint64_t seed = (EPI == B200U_EPI_BIAS_DROP_RES) ? load_seed(g.drop) : 0ull;
    epilogue_row32<EPI>(acc, row, col0, g, seed);
}

// ---------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------
extern long long* g_attn_dbg;  // attention_tc.cu
static long long* g_dbg_ptr = nullptr;
static int g_dbg_mode = 0;

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, int CLUSTER>
static int launch_tc(const b200u_gemm_t* d, GemmArgs& g, cudaStream_t stream) {
    using Cfg = GemmCfg<BLOCK_N, EPI>;
    CUtensorMap tmA, tmB, tmC, tmC2, tmR;
    int rc;
    if (!A_MN) rc = make_tmap(&tmA, d->A, d->M, d->K, d->lda, BLOCK_M);
    else       rc = make_tmap(&tmA, d->A, d->K, d->M, d->lda, BLOCK_K);
    if (rc) return rc;
    if (!B_MN) rc = make_tmap(&tmB, d->B, d->N, d->K, d->ldb, BLOCK_N / CLUSTER);
    else       rc = make_tmap(&tmB, d->B, d->K, d->N, d->ldb, BLOCK_K);
    if (rc) return rc;
    if (Cfg::CE_STATS) tmC = tmA;   // no output matrix: the epilogue writes row statistics
    else rc = make_tmap(&tmC, d->C, d->M, d->N, d->ldc, 32, Cfg::F32_OUT);
    if (rc) return rc;
    tmC2 = tmC;
    tmR = tmC;
    if (Cfg::DUAL || Cfg::LN) { rc = make_tmap(&tmC2, d->C2, d->M, d->N, d->ldc2, 32); if (rc) return rc; }
    if (Cfg::HAS_R) { rc = make_tmap(&tmR, d->R, d->M, d->N, d->ldr, BLOCK_M); if (rc) return rc; }

    g.m_tiles = (d->M + BLOCK_M - 1) / BLOCK_M;
    g.n_tiles = (d->N + BLOCK_N - 1) / BLOCK_N;
    const int ucl = Cfg::LN ? g.ln_cl : CLUSTER;  // cluster size of the launch
    const int units = Cfg::LN ? g.m_tiles : ((g.m_tiles + CLUSTER - 1) / CLUSTER) * g.n_tiles * g.splits;

    auto kern = gemm_tc_kernel<BLOCK_N, A_MN, B_MN, EPI, CLUSTER>;
    // per-device, per-instantiation one-time setup (the attribute is per device; PyTorch calls us from
    // the main and the autograd thread)
    static std::mutex mu;
    static bool attr_set[64] = {};
    static int ln_max_clusters[64][Cfg::LN_MAX_CL + 1] = {};
    int dev = 0;
    B200U_CHECK_CUDA(cudaGetDevice(&dev));
    B200U_CHECK_ARG(dev >= 0 && dev < 64, "b200u_gemm: device ordinal %d out of range", dev);
    int max_clusters = num_sms() / ucl;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!attr_set[dev]) {
            B200U_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  Cfg::SMEM_BYTES));
            attr_set[dev] = true;
        }
        if (Cfg::LN) {
            // how many clusters of this size the device can hold at once (GPC boundaries): the persistent
            // loop must not be sized past it, or the surplus clusters would only start after a full pass
            if (ln_max_clusters[dev][ucl] == 0) {
                cudaLaunchConfig_t q = {};
                q.gridDim = dim3(ucl * (num_sms() / ucl));
                q.blockDim = dim3(GEMM_THREADS);
                q.dynamicSmemBytes = Cfg::SMEM_BYTES;
                cudaLaunchAttribute qa[1];
                qa[0].id = cudaLaunchAttributeClusterDimension;
                qa[0].val.clusterDim.x = ucl; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                q.attrs = qa; q.numAttrs = 1;
                int n = 0;
                if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n < 1) {
                    cudaGetLastError();
                    n = 1;
                }
                ln_max_clusters[dev][ucl] = n;
            }
            if (ln_max_clusters[dev][ucl] < max_clusters) max_clusters = ln_max_clusters[dev][ucl];
        }
    }
    if (max_clusters < 1) max_clusters = 1;
    const int grid = (units < max_clusters ? units : max_clusters) * ucl;
    int slot = 0;
    const bool prof = prof_begin(stream, 2.0 * d->M * d->N * d->K, &slot);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (ucl > 1 || Cfg::LN) {  // (LN with one column tile: an explicit 1-CTA cluster, st.async needs a cluster launch)
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = ucl;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    B200U_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmC2, tmR, g));
    if (prof) prof_end(stream, slot);
    B200U_CHECK_LAUNCH("gemm_tc_kernel");
    return B200U_OK;
}

template <int BLOCK_N, int EPI, int CLUSTER>
static int dispatch_major(const b200u_gemm_t* d, GemmArgs& g, cudaStream_t stream) {
    if (!d->a_mn_major && !d->b_mn_major) return launch_tc<BLOCK_N, false, false, EPI, CLUSTER>(d, g, stream);
    if (!d->a_mn_major && d->b_mn_major) return launch_tc<BLOCK_N, false, true, EPI, CLUSTER>(d, g, stream);
    if (d->a_mn_major && d->b_mn_major) return launch_tc<BLOCK_N, true, true, EPI, CLUSTER>(d, g, stream);
    return launch_tc<BLOCK_N, true, false, EPI, CLUSTER>(d, g, stream);
}

template <int EPI>
static int dispatch_bn(const b200u_gemm_t* d, GemmArgs& g, int block_n, cudaStream_t stream) {
    if constexpr (EPI == B200U_EPI_CE_STATS || EPI == B200U_EPI_CE_GRAD) {
        // hidden[M,K] . W[N,K]^T only, tcgen05 path only. Statistics are indexed by 128-column tiles.
        B200U_CHECK_ARG(d->impl == 0 && !d->a_mn_major && !d->b_mn_major,
                        "b200u_gemm: the cross-entropy epilogues need K-major A and B on the tcgen05 path");
        if (EPI == B200U_EPI_CE_GRAD && block_n == 256) return launch_tc<256, false, false, EPI, 1>(d, g, stream);
        return launch_tc<128, false, false, EPI, 1>(d, g, stream);
    } else
    if (d->impl == 1) {
        dim3 grid((d->N + 31) / 32, (d->M + 127) / 128);
        g.splits = 1;
        constexpr int SEPI = EPI == B200U_EPI_BIAS_DROP_RES_LN ? B200U_EPI_BIAS_DROP_RES : EPI;
        launch_k(gemm_simt_kernel<SEPI>, dim3(grid), dim3(128), 0, stream, (const bf16*)d->A, d->lda, d->a_mn_major,
                                                        (const bf16*)d->B, d->ldb, d->b_mn_major, g);
        B200U_CHECK_LAUNCH("gemm_simt_kernel");
        if (EPI == B200U_EPI_BIAS_DROP_RES_LN) {
            B200U_CHECK_ARG(d->ldc == d->N && d->ldc2 == d->N, "b200u_gemm: SIMT LayerNorm epilogue needs dense C / C2");
            return b200u_layernorm_fwd(d->C, B200U_BF16, d->ln_gamma, d->ln_beta, d->C2, B200U_BF16, d->ln_mean,
                                       d->ln_rstd, d->M, d->N, d->ln_eps, nullptr, (b200u_stream_t)stream);
        }
        return B200U_OK;
    }
    if constexpr (EPI == B200U_EPI_BIAS_DROP_RES_LN) {
        // forward-only epilogue: X[M,K] . W[N,K]^T, one cluster of N / tile-width CTAs per 128-row block.
        // 128-wide tiles (cluster of N/128 <= 8) keep more SMs busy; 256-wide tiles (cluster of N/256 <= 3)
        // halve the operand ingest per FLOP and the number of clusters: they win once the narrow
        // configuration would need a second pass over the row blocks.
        B200U_CHECK_ARG(!d->a_mn_major && !d->b_mn_major, "b200u_gemm: the LayerNorm epilogue needs K-major A and B");
        const int m_tiles = (d->M + BLOCK_M - 1) / BLOCK_M;
        const bool can128 = d->N % 128 == 0 && d->N / 128 <= 8;
        const bool can256 = d->N % 256 == 0 && d->N / 256 <= 3;
        B200U_CHECK_ARG(can128 || can256, "b200u_gemm: the LayerNorm epilogue needs N = 128 k (k <= 8) or 256 k (k <= 3), got %d", d->N);
        bool wide = !can128;
        if (can128 && can256) {
            const int cap128 = num_sms() / (d->N / 128);   // upper bound of co-resident clusters
            wide = block_n == 256 || (block_n == 0 && m_tiles > cap128);
        }
        if (wide) {
            g.ln_cl = d->N / 256;
            return launch_tc<256, false, false, EPI, 1>(d, g, stream);
        }
        g.ln_cl = d->N / 128;
        return launch_tc<128, false, false, EPI, 1>(d, g, stream);
    } else {
        // pairs of vertically adjacent tiles share their B tile through TMA multicast
        // (measured on B200: multicast at cluster size 2 does not reduce per-SM ingest, so auto = off)
        const bool pair = d->cluster == 2 && EPI != B200U_EPI_BIAS_GELU_DG && EPI != B200U_EPI_MUL;
        if constexpr (EPI != B200U_EPI_BIAS_GELU_DG && EPI != B200U_EPI_MUL) {
            if (pair) {
                if (block_n == 256) return dispatch_major<256, EPI, 2>(d, g, stream);
                return dispatch_major<128, EPI, 2>(d, g, stream);
            }
        }
        if (block_n == 256) return dispatch_major<256, EPI, 1>(d, g, stream);
        return dispatch_major<128, EPI, 1>(d, g, stream);
    }
}

// one warp per row: merge the (max, sum) partials of the row's column tiles
__global__ void __launch_bounds__(256)
ce_finish_kernel(const float2* __restrict__ partial, const float* __restrict__ tlogit, float* __restrict__ lse,
                 float* __restrict__ loss, int M, int n_tiles) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    float m = -INFINITY, s_ = 0.f;
    for (int t = lane; t < n_tiles; t += 32) {
        const float2 p = partial[(size_t)row * n_tiles + t];
        const float mm = fmaxf(m, p.x);
        s_ = (m > -INFINITY ? s_ * __expf(m - mm) : 0.f) + (p.x > -INFINITY ? p.y * __expf(p.x - mm) : 0.f);
        m = mm;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s_, o);
        const float mm = fmaxf(m, m2);
        s_ = (m > -INFINITY ? s_ * __expf(m - mm) : 0.f) + (m2 > -INFINITY ? s2 * __expf(m2 - mm) : 0.f);
        m = mm;
    }
    if (lane == 0) {
        const float l = m + logf(s_);
        lse[row] = l;
        if (loss) loss[row] = l - tlogit[row];
    }
}

}  // namespace b200u

using namespace b200u;

extern "C" int b200u_ce_finish(const float* partial, const float* tlogit, float* lse, float* loss, int M, int n_tiles,
                               b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(partial && tlogit && lse && n_tiles > 0 && ((uintptr_t)partial & 7) == 0, "ce_finish: bad arguments");
    if (M == 0) return B200U_OK;
    launch_k(ce_finish_kernel, dim3((M + 7) / 8), dim3(256), 0, stream, (const float2*)partial, tlogit, lse, loss, M, n_tiles);
    B200U_CHECK_LAUNCH("ce_finish");
    return B200U_OK;
}

// Bring-up aid: when set, every tcgen05 GEMM CTA writes 8 clock64() phase stamps to ptr[cta*8..].
extern "C" int b200u_gemm_debug_stamps(long long* device_ptr) {
    // low 2 bits of the (8-byte aligned) pointer select a bring-up mode: 1 = TMA-only, 2 = MMA-only
    g_dbg_mode = (int)((uintptr_t)device_ptr & 3);
    device_ptr = (long long*)((uintptr_t)device_ptr & ~(uintptr_t)3);
    g_dbg_ptr = device_ptr;
    g_attn_dbg = device_ptr;
    return B200U_OK;
}

extern "C" int b200u_gemm(const b200u_gemm_t* d, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(d != nullptr, "b200u_gemm: null descriptor");
    B200U_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "b200u_gemm: bad shape M=%d N=%d K=%d", d->M,
                    d->N, d->K);
    const bool ce_stats = d->epilogue == B200U_EPI_CE_STATS, ce_grad = d->epilogue == B200U_EPI_CE_GRAD;
    B200U_CHECK_ARG(d->A && d->B && (d->C || ce_stats), "b200u_gemm: null operand pointer");
    if (ce_stats)
        B200U_CHECK_ARG(d->ce_target && d->ce_partial && d->ce_tlogit && d->bias && ((uintptr_t)d->ce_partial & 7) == 0,
                        "b200u_gemm: EPI_CE_STATS needs bias, ce_target, ce_partial and ce_tlogit");
    if (ce_grad)
        B200U_CHECK_ARG(d->ce_target && d->ce_lse && d->ce_scale && d->bias,
                        "b200u_gemm: EPI_CE_GRAD needs bias, ce_target, ce_lse and ce_scale");
    B200U_CHECK_ARG(d->epilogue >= 0 && d->epilogue < B200U_EPI_COUNT, "b200u_gemm: bad epilogue %d",
                    d->epilogue);
    B200U_CHECK_ARG(d->lda % 8 == 0 && d->ldb % 8 == 0,
                    "b200u_gemm: lda/ldb must be multiples of 8 elements (TMA 16-byte strides)");
    B200U_CHECK_ARG(((uintptr_t)d->A & 15) == 0 && ((uintptr_t)d->B & 15) == 0 &&
                        ((uintptr_t)d->C & 15) == 0,
                    "b200u_gemm: operands must be 16-byte aligned");
    const bool f32_out = d->epilogue == B200U_EPI_ATOMIC_F32 || d->epilogue == B200U_EPI_STORE_F32;
    B200U_CHECK_ARG(d->ldc % (f32_out ? 4 : 8) == 0,
                    "b200u_gemm: ldc alignment (got %d)", d->ldc);
    if (d->epilogue == B200U_EPI_BIAS_GELU || d->epilogue == B200U_EPI_BIAS_GELU_DG)
        B200U_CHECK_ARG(d->C2 && d->bias && d->ldc2 % 8 == 0, "b200u_gemm: BIAS_GELU needs C2 and bias");
    const bool ln = d->epilogue == B200U_EPI_BIAS_DROP_RES_LN;
    if (ln) {
        B200U_CHECK_ARG(d->C2 && d->ldc2 % 8 == 0 && ((uintptr_t)d->C2 & 15) == 0 && d->ln_gamma && d->ln_beta && d->bias,
                        "b200u_gemm: the LayerNorm epilogue needs C2, bias, ln_gamma and ln_beta");
    }
    if (d->epilogue == B200U_EPI_BIAS_DROP_RES || d->epilogue == B200U_EPI_ADD ||
        d->epilogue == B200U_EPI_DGELU || d->epilogue == B200U_EPI_MUL || ln)
        B200U_CHECK_ARG(d->R && d->ldr % 8 == 0 && ((uintptr_t)d->R & 15) == 0,
                        "b200u_gemm: epilogue %d needs R", d->epilogue);
    B200U_CHECK_ARG(d->splits <= 1 || d->epilogue == B200U_EPI_ATOMIC_F32,
                    "b200u_gemm: split-K requires EPI_ATOMIC_F32");
    B200U_CHECK_ARG(d->block_n == 0 || d->block_n == 128 || d->block_n == 256,
                    "b200u_gemm: block_n must be 0, 128 or 256");
    B200U_CHECK_ARG(!d->colsum || d->epilogue == B200U_EPI_MUL, "b200u_gemm: colsum is an EPI_MUL output");

    GemmArgs g;
    g.M = d->M; g.N = d->N; g.K = d->K;
    g.C = d->C; g.ldc = d->ldc; g.C2 = d->C2; g.ldc2 = d->ldc2;
    g.bias = d->bias; g.R = (const bf16*)d->R; g.ldr = d->ldr;
    g.drop.seed_ptr = d->drop.seed_ptr;
    g.drop.stream = d->drop.stream;
    g.colsum = d->colsum;
    g.ln_gamma = d->ln_gamma; g.ln_beta = d->ln_beta; g.ln_mean = d->ln_mean; g.ln_rstd = d->ln_rstd;
    g.ln_eps = d->ln_eps;
    g.ln_cl = 1;  // set by dispatch (N / tile width)
    g.ce_target = d->ce_target; g.ce_partial = d->ce_partial; g.ce_tlogit = d->ce_tlogit;
    g.ce_lse = d->ce_lse; g.ce_scale = d->ce_scale;
    const float p = (d->epilogue == B200U_EPI_BIAS_DROP_RES || ln) ? d->drop.p : 0.f;
    B200U_CHECK_ARG(p >= 0.f && p < 1.f, "b200u_gemm: dropout p out of range");
    g.drop.thresh16 = (uint32_t)(p * 65536.0f + 0.5f);
    g.drop.scale = 1.0f / (1.0f - p);
    B200U_CHECK_ARG(g.drop.thresh16 == 0 || g.drop.seed_ptr, "b200u_gemm: dropout needs seed_ptr");
    g.num_kb = (d->K + BLOCK_K - 1) / BLOCK_K;

    // tile shape: wide tiles when they still fill the machine, else 128. The fp32 reduce-add
    // (wgrad) family always prefers 256-wide tiles (MN-major x MN-major 128-wide tiles are shared
    // memory bandwidth bound) and as many K-splits as still fit ONE wave of CTAs -- measured on
    // B200 at the C2 shapes (tools/gemm_sweep.py): 768x3072x2624 bn256/s2 14.3 us vs bn128/s1 17.4.
    int block_n = d->block_n;
    const int m_tiles = (d->M + BLOCK_M - 1) / BLOCK_M;
    const bool reduce = d->epilogue == B200U_EPI_ATOMIC_F32;
    if (block_n == 0) {
        // cost model from the measured main-loop rates (cycles per 64-deep k-block: ~420 at 128-wide tiles,
        // where the per-SM operand ingest is the limit, ~550 at 256-wide tiles, where the MMAs are) times
        // the number of tiles the busiest CTA walks: e.g. N = 768 at M = 5248 is one wave of 123 wide tiles
        // (550 per k-block) rather than two waves of 246 narrow ones (2 x 420)
        const int sms = num_sms();
        const int t128 = m_tiles * ((d->N + 127) / 128), t256 = m_tiles * ((d->N + 255) / 256);
        const long c128 = (long)((t128 + sms - 1) / sms) * 420, c256 = (long)((t256 + sms - 1) / sms) * 550;
        if (reduce) block_n = d->N >= 256 ? 256 : 128;
        else block_n = (d->N >= 256 && c256 <= c128) ? 256 : 128;
    }
    if (ce_stats) block_n = 128;
    // (epilogues with a side-input tile keep it in smem and write their output over it in place:
    //  3 pipeline stages remain at 256-wide tiles, 5 at 128)
    int splits = d->splits;
    if (!reduce) splits = 1;
    else if (splits <= 0) {
        const int tiles = m_tiles * ((d->N + block_n - 1) / block_n);
        splits = num_sms() / tiles;
        if (splits > g.num_kb) splits = g.num_kb;
        if (splits < 1) splits = 1;
    }
    g.kb_per_split = (g.num_kb + splits - 1) / splits;
    g.splits = (g.num_kb + g.kb_per_split - 1) / g.kb_per_split;
    g.m_tiles = m_tiles;
    g.n_tiles = 0;
    g.dbg = g_dbg_ptr;
    g.dbg_mode = g_dbg_mode;

    switch (d->epilogue) {
        case B200U_EPI_STORE:         return dispatch_bn<B200U_EPI_STORE>(d, g, block_n, stream);
        case B200U_EPI_BIAS_GELU:     return dispatch_bn<B200U_EPI_BIAS_GELU>(d, g, block_n, stream);
        case B200U_EPI_BIAS_DROP_RES: return dispatch_bn<B200U_EPI_BIAS_DROP_RES>(d, g, block_n, stream);
        case B200U_EPI_ADD:           return dispatch_bn<B200U_EPI_ADD>(d, g, block_n, stream);
        case B200U_EPI_DGELU:         return dispatch_bn<B200U_EPI_DGELU>(d, g, block_n, stream);
        case B200U_EPI_ATOMIC_F32:    return dispatch_bn<B200U_EPI_ATOMIC_F32>(d, g, block_n, stream);
        case B200U_EPI_STORE_F32:     return dispatch_bn<B200U_EPI_STORE_F32>(d, g, block_n, stream);
        case B200U_EPI_BIAS_GELU_DG:  return dispatch_bn<B200U_EPI_BIAS_GELU_DG>(d, g, block_n, stream);
        case B200U_EPI_MUL:           return dispatch_bn<B200U_EPI_MUL>(d, g, block_n, stream);
        case B200U_EPI_BIAS_DROP_RES_LN: return dispatch_bn<B200U_EPI_BIAS_DROP_RES_LN>(d, g, d->block_n, stream);
        case B200U_EPI_CE_STATS:      return dispatch_bn<B200U_EPI_CE_STATS>(d, g, block_n, stream);
        case B200U_EPI_CE_GRAD:       return dispatch_bn<B200U_EPI_CE_GRAD>(d, g, block_n, stream);
    }
    return B200U_ERR_ARG;
}
