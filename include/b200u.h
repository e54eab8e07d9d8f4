/*
 * b200u — C-ABI of the B200-native UNITER hot path (sm_100a).
 *
 * The reference (Nithin-Holla/meme_challenge) has no FFI layer of its own: the hot path sits
 * behind Python nn.Module calls. This header is therefore the boundary the
 * Python mirror in meme_challenge_b200/ binds with ctypes; each entry point cites the reference
 * lines whose arithmetic it replaces. INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless marked host.
 *  - the caller (PyTorch) owns every buffer; the library never allocates, frees or retains.
 *  - every function enqueues on `stream` and returns immediately (CUDA-graph capturable):
 *    0 on success, negative on error; b200u_last_error_string() describes the last error
 *    of the calling thread. No exceptions, no exit().
 *  - bf16 = __nv_bfloat16 storage; "f32" = float. Row-major everywhere, `ld*` in elements.
 */
#ifndef B200U_H_
#define B200U_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200u_stream_t; /* cudaStream_t */

const char* b200u_last_error_string(void);
int b200u_version(void);
/* Compiled SASS arch (100 for sm_100a) and SM count of the current device. */
int b200u_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * Dropout configuration shared by all fused-dropout kernels (torch nn.Dropout sites:
 * model/model.py:230,259; model/layer.py:68,109,150). seed_ptr is a DEVICE uint64 so graph
 * replays draw fresh masks; p == 0 (eval mode) disables the site. */
typedef struct {
    const unsigned long long* seed_ptr;
    uint32_t stream; /* unique id per (layer, site) */
    float p;
} b200u_dropout_t;

/* ------------------------------------------------------------------------------------------
 * K3  GEMM on tcgen05 tensor cores fed by TMA  (every nn.Linear on the path:
 * model/layer.py:64-66,107,133,148,176; model/model.py:250,253; and their autograd backward).
 *
 *   acc[M,N] = sum_k A(m,k) * B(n,k)     bf16 operands, fp32 accumulate in TMEM
 *
 * a_mn_major = 0: A is stored [M,K] (lda = row stride);  1: A is stored [K,M] (transposed use,
 * wgrad dY^T).  b_mn_major = 0: B is stored [N,K] (nn.Linear weight);  1: B is stored [K,N]
 * (dgrad reads the same weight without a transposed copy). */
enum {
    B200U_EPI_STORE = 0,         /* C(bf16) = acc (+ bias[n] if bias)                         */
    B200U_EPI_BIAS_GELU = 1,     /* C(bf16) = u = acc + bias ; C2(bf16) = gelu_erf(u)          */
    B200U_EPI_BIAS_DROP_RES = 2, /* C(bf16) = dropout(acc + bias) + R                         */
    B200U_EPI_ADD = 3,           /* C(bf16) = acc + R                                         */
    B200U_EPI_DGELU = 4,         /* C(bf16) = acc * gelu_erf'(R)                              */
    B200U_EPI_ATOMIC_F32 = 5,    /* C(f32) += acc   (split-K safe; wgrad into .grad buffers)  */
    B200U_EPI_STORE_F32 = 6,     /* C(f32) = acc (+ bias)                                     */
    B200U_EPI_COUNT = 7
};

typedef struct {
    int M, N, K;
    const void* A; int lda; int a_mn_major;
    const void* B; int ldb; int b_mn_major;
    int epilogue;
    void* C; int ldc;
    void* C2; int ldc2;
    const float* bias;     /* f32 [N] or NULL */
    const void* R; int ldr; /* bf16 [M,N] side input (residual / pre-activation) or NULL */
    b200u_dropout_t drop;   /* EPI_BIAS_DROP_RES only */
    int splits;             /* split-K factor, >1 only with EPI_ATOMIC_F32; 0 = auto */
    int block_n;            /* 0 = auto, else 128 or 256 */
    int impl;               /* 0 = tcgen05 (product path), 1 = SIMT debug kernel */
} b200u_gemm_t;

int b200u_gemm(const b200u_gemm_t* g, b200u_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200U_H_ */
