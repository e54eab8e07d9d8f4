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
};

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
            EPI == B200U_EPI_STORE_F32) {
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
        if (EPI == B200U_EPI_BIAS_DROP_RES || EPI == B200U_EPI_ADD || EPI == B200U_EPI_DGELU) {
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

// smem matrix descriptor, SWIZZLE_128B, version 1 (Blackwell).
//  K-major : rows of 128 B (64 k-elements); 8-row groups SBO = 1024 B apart; LBO unused.
//  MN-major: rows of 128 B (64 mn-elements), one row per k; 8-k groups SBO = 1024 B apart;
//            next 64-wide mn block LBO bytes away (= one TMA box = BLOCK_K * 128 B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
// instruction descriptor for kind::f16, bf16 x bf16 -> f32.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn) {
    return (1u << 4)                      // D format f32
           | (1u << 7)                    // A bf16
           | (1u << 10)                   // B bf16
           | ((a_mn ? 1u : 0u) << 15)     // A major
           | ((b_mn ? 1u : 0u) << 16)     // B major
           | ((uint32_t)(n >> 3) << 17)   // N
           | ((uint32_t)(m >> 4) << 24);  // M
}

template <int BLOCK_N>
struct GemmCfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGES = (BLOCK_N >= 256) ? 4 : 6;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages
    static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*align*/ + 256 /*bars*/;
};

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmArgs g) {
    using Cfg = GemmCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_mn = g.m_tiles * g.n_tiles;
    const int total_tiles = tiles_mn * g.splits;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int split = t / tiles_mn;
                const int rem = t - split * tiles_mn;
                const int m0 = (rem % g.m_tiles) * BLOCK_M;
                const int n0 = (rem / g.m_tiles) * BLOCK_N;
                const int kb0 = split * g.kb_per_split;
                const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], Cfg::A_BYTES + Cfg::B_BYTES);
                    uint8_t* a_dst = sA + s * Cfg::A_BYTES;
                    uint8_t* b_dst = sB + s * Cfg::B_BYTES;
                    if (!A_MN) {
                        tma_load_2d(a_dst, &tmA, &full[s], kb * BLOCK_K, m0);
                    } else {
#pragma unroll
                        for (int i = 0; i < BLOCK_M / 64; ++i)
                            tma_load_2d(a_dst + i * (BLOCK_K * 128), &tmA, &full[s], m0 + 64 * i,
                                        kb * BLOCK_K);
                    }
                    if (!B_MN) {
                        tma_load_2d(b_dst, &tmB, &full[s], kb * BLOCK_K, n0);
                    } else {
#pragma unroll
                        for (int i = 0; i < BLOCK_N / 64; ++i)
                            tma_load_2d(b_dst + i * (BLOCK_K * 128), &tmB, &full[s], n0 + 64 * i,
                                        kb * BLOCK_K);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, A_MN, B_MN);
        int s = 0;
        uint32_t ph = 0;
        int as = 0;
        uint32_t aph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int split = t / tiles_mn;
            const int kb0 = split * g.kb_per_split;
            const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
            mbar_wait(&tempty[as], aph ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * BLOCK_N;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(sA + s * Cfg::A_BYTES);
                    const uint32_t b_addr = smem_u32(sB + s * Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t ad =
                            A_MN ? make_smem_desc(a_addr + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                 : make_smem_desc(a_addr + k * (UMMA_K * 2), 16, 1024);
                        const uint64_t bd =
                            B_MN ? make_smem_desc(b_addr + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                 : make_smem_desc(b_addr + k * (UMMA_K * 2), 16, 1024);
                        umma_bf16(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);
                    if (kb == kb1 - 1) umma_commit(&tfull[as]);
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            as ^= 1;
            if (as == 0) aph ^= 1;
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        const uint64_t seed = (EPI == B200U_EPI_BIAS_DROP_RES) ? load_seed(g.drop) : 0ull;
        int as = 0;
        uint32_t aph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int split = t / tiles_mn;
            const int rem = t - split * tiles_mn;
            const int m0 = (rem % g.m_tiles) * BLOCK_M;
            const int n0 = (rem / g.m_tiles) * BLOCK_N;
            mbar_wait(&tfull[as], aph);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c * 32, r);
                tmem_ld_wait();
                float acc[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = __uint_as_float(r[i]);
                epilogue_row32<EPI>(acc, row, n0 + c * 32, g, seed);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            as ^= 1;
            if (as == 0) aph ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------
// SIMT debug kernel: same contract and the same epilogue, no tensor cores. Selected only by
// b200u_gemm_t.impl = 1 (bring-up / differential testing of the tcgen05 path on the GPU).
// ---------------------------------------------------------------------------------------
template <int EPI>
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, int lda, int a_mn,
                                 const bf16* __restrict__ B, int ldb, int b_mn, const GemmArgs g) {
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
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// bf16 matrix stored as [rows, cols] row-major with leading dimension ld; box = 64 cols x box_rows.
static int make_tmap(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld, int box_rows) {
    auto fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return B200U_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%d cols=%d ld=%d box_rows=%d", (int)r,
                  ptr, rows, cols, ld, box_rows);
        return B200U_ERR_CUDA;
    }
    return B200U_OK;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
static int launch_tc(const b200u_gemm_t* d, GemmArgs& g, cudaStream_t stream) {
    using Cfg = GemmCfg<BLOCK_N>;
    CUtensorMap tmA, tmB;
    int rc;
    if (!A_MN) rc = make_tmap(&tmA, d->A, d->M, d->K, d->lda, BLOCK_M);
    else       rc = make_tmap(&tmA, d->A, d->K, d->M, d->lda, BLOCK_K);
    if (rc) return rc;
    if (!B_MN) rc = make_tmap(&tmB, d->B, d->N, d->K, d->ldb, BLOCK_N);
    else       rc = make_tmap(&tmB, d->B, d->K, d->N, d->ldb, BLOCK_K);
    if (rc) return rc;

    g.m_tiles = (d->M + BLOCK_M - 1) / BLOCK_M;
    g.n_tiles = (d->N + BLOCK_N - 1) / BLOCK_N;
    const int total = g.m_tiles * g.n_tiles * g.splits;
    const int grid = total < num_sms() ? total : num_sms();

    auto kern = gemm_tc_kernel<BLOCK_N, A_MN, B_MN, EPI>;
    static bool attr_set = false;  // per template instantiation
    if (!attr_set) {
        B200U_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              Cfg::SMEM_BYTES));
        attr_set = true;
    }
    int slot = 0;
    const bool prof = prof_begin(stream, 2.0 * d->M * d->N * d->K, &slot);
    kern<<<grid, 256, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, g);
    if (prof) prof_end(stream, slot);
    B200U_CHECK_LAUNCH("gemm_tc_kernel");
    return B200U_OK;
}

template <int BLOCK_N, int EPI>
static int dispatch_major(const b200u_gemm_t* d, GemmArgs& g, cudaStream_t stream) {
    if (!d->a_mn_major && !d->b_mn_major) return launch_tc<BLOCK_N, false, false, EPI>(d, g, stream);
    if (!d->a_mn_major && d->b_mn_major) return launch_tc<BLOCK_N, false, true, EPI>(d, g, stream);
    if (d->a_mn_major && d->b_mn_major) return launch_tc<BLOCK_N, true, true, EPI>(d, g, stream);
    return launch_tc<BLOCK_N, true, false, EPI>(d, g, stream);
}

template <int EPI>
static int dispatch_bn(const b200u_gemm_t* d, GemmArgs& g, int block_n, cudaStream_t stream) {
    if (d->impl == 1) {
        dim3 grid((d->N + 31) / 32, (d->M + 127) / 128);
        g.splits = 1;
        gemm_simt_kernel<EPI><<<grid, 128, 0, stream>>>((const bf16*)d->A, d->lda, d->a_mn_major,
                                                        (const bf16*)d->B, d->ldb, d->b_mn_major, g);
        B200U_CHECK_LAUNCH("gemm_simt_kernel");
        return B200U_OK;
    }
    if (block_n == 256) return dispatch_major<256, EPI>(d, g, stream);
    return dispatch_major<128, EPI>(d, g, stream);
}

}  // namespace b200u

using namespace b200u;

extern "C" int b200u_gemm(const b200u_gemm_t* d, b200u_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200U_CHECK_ARG(d != nullptr, "b200u_gemm: null descriptor");
    B200U_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "b200u_gemm: bad shape M=%d N=%d K=%d", d->M,
                    d->N, d->K);
    B200U_CHECK_ARG(d->A && d->B && d->C, "b200u_gemm: null operand pointer");
    B200U_CHECK_ARG(d->epilogue >= 0 && d->epilogue < B200U_EPI_COUNT, "b200u_gemm: bad epilogue %d",
                    d->epilogue);
    B200U_CHECK_ARG(d->lda % 8 == 0 && d->ldb % 8 == 0,
                    "b200u_gemm: lda/ldb must be multiples of 8 elements (TMA 16-byte strides)");
    B200U_CHECK_ARG(((uintptr_t)d->A & 15) == 0 && ((uintptr_t)d->B & 15) == 0 &&
                        ((uintptr_t)d->C & 15) == 0,
                    "b200u_gemm: operands must be 16-byte aligned");
    const bool f32_out = d->epilogue == B200U_EPI_ATOMIC_F32 || d->epilogue == B200U_EPI_STORE_F32;
    B200U_CHECK_ARG(d->ldc % (f32_out ? 4 : 8) == 0 || d->N < 8,
                    "b200u_gemm: ldc alignment (got %d)", d->ldc);
    if (d->epilogue == B200U_EPI_BIAS_GELU)
        B200U_CHECK_ARG(d->C2 && d->bias && d->ldc2 % 8 == 0, "b200u_gemm: BIAS_GELU needs C2 and bias");
    if (d->epilogue == B200U_EPI_BIAS_DROP_RES || d->epilogue == B200U_EPI_ADD ||
        d->epilogue == B200U_EPI_DGELU)
        B200U_CHECK_ARG(d->R && d->ldr % 8 == 0 && ((uintptr_t)d->R & 15) == 0,
                        "b200u_gemm: epilogue %d needs R", d->epilogue);
    B200U_CHECK_ARG(d->splits <= 1 || d->epilogue == B200U_EPI_ATOMIC_F32,
                    "b200u_gemm: split-K requires EPI_ATOMIC_F32");
    B200U_CHECK_ARG(d->block_n == 0 || d->block_n == 128 || d->block_n == 256,
                    "b200u_gemm: block_n must be 0, 128 or 256");

    GemmArgs g;
    g.M = d->M; g.N = d->N; g.K = d->K;
    g.C = d->C; g.ldc = d->ldc; g.C2 = d->C2; g.ldc2 = d->ldc2;
    g.bias = d->bias; g.R = (const bf16*)d->R; g.ldr = d->ldr;
    g.drop.seed_ptr = d->drop.seed_ptr;
    g.drop.stream = d->drop.stream;
    const float p = (d->epilogue == B200U_EPI_BIAS_DROP_RES) ? d->drop.p : 0.f;
    B200U_CHECK_ARG(p >= 0.f && p < 1.f, "b200u_gemm: dropout p out of range");
    g.drop.thresh16 = (uint32_t)(p * 65536.0f + 0.5f);
    g.drop.scale = 1.0f / (1.0f - p);
    B200U_CHECK_ARG(g.drop.thresh16 == 0 || g.drop.seed_ptr, "b200u_gemm: dropout needs seed_ptr");
    g.num_kb = (d->K + BLOCK_K - 1) / BLOCK_K;

    // tile shape: wide tiles when they still fill the machine, else 128.
    int block_n = d->block_n;
    const int m_tiles = (d->M + BLOCK_M - 1) / BLOCK_M;
    if (block_n == 0) {
        const int t256 = m_tiles * ((d->N + 255) / 256);
        block_n = (d->N >= 256 && t256 >= num_sms()) ? 256 : 128;
    }
    int splits = d->splits;
    if (d->epilogue != B200U_EPI_ATOMIC_F32) splits = 1;
    else if (splits <= 0) {
        const int tiles = m_tiles * ((d->N + block_n - 1) / block_n);
        splits = (num_sms() + tiles - 1) / tiles;
        if (splits > g.num_kb) splits = g.num_kb;
        if (splits < 1) splits = 1;
    }
    g.kb_per_split = (g.num_kb + splits - 1) / splits;
    g.splits = (g.num_kb + g.kb_per_split - 1) / g.kb_per_split;
    g.m_tiles = m_tiles;
    g.n_tiles = 0;

    switch (d->epilogue) {
        case B200U_EPI_STORE:         return dispatch_bn<B200U_EPI_STORE>(d, g, block_n, stream);
        case B200U_EPI_BIAS_GELU:     return dispatch_bn<B200U_EPI_BIAS_GELU>(d, g, block_n, stream);
        case B200U_EPI_BIAS_DROP_RES: return dispatch_bn<B200U_EPI_BIAS_DROP_RES>(d, g, block_n, stream);
        case B200U_EPI_ADD:           return dispatch_bn<B200U_EPI_ADD>(d, g, block_n, stream);
        case B200U_EPI_DGELU:         return dispatch_bn<B200U_EPI_DGELU>(d, g, block_n, stream);
        case B200U_EPI_ATOMIC_F32:    return dispatch_bn<B200U_EPI_ATOMIC_F32>(d, g, block_n, stream);
        case B200U_EPI_STORE_F32:     return dispatch_bn<B200U_EPI_STORE_F32>(d, g, block_n, stream);
    }
    return B200U_ERR_ARG;
}
