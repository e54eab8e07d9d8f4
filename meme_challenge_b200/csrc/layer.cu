// BertLayer forward / backward as one C call each (model/layer.py:159-170): the launcher strings
// together the tcgen05 GEMMs, the fused attention and the LayerNorm kernels so the Python side
// pays one ctypes call per layer and the whole sequence is CUDA-graph capturable.
//
// forward (5 launches)                                   reference
//   qkv = x0·Wqkvᵀ + bqkv                                layer.py:76-78   (one N=3H GEMM)
//   ctx = attention(qkv, mask)                           layer.py:80-100
//   y1  = dropout(ctx·Woᵀ + bo) + x0 ; x1 = LN1(y1)      layer.py:111-115 (ONE GEMM: LayerNorm in the epilogue,
//   u   = x1·W1ᵀ + b1 ; g = gelu(u) ; saves gelu'(u)     layer.py:139-142  row statistics exchanged over DSMEM
//   y2  = dropout(g·W2ᵀ + b2) + x1 ; x2 = LN2(y2)        layer.py:152-156  by the cluster that owns a row block)
// (H not a multiple of 128 or > 1024: LayerNorm runs as its own launch, 7 launches.)
// backward (11 launches): the exact transposes, weight grads accumulated in fp32 (+=); the FFN1 bias
// gradient is a by-product of the FFN2 dgrad epilogue (du = (dz2·W2) * gelu'(u), column sums folded in).
#include "../../include/b200u.h"
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

#include <mutex>

using namespace b200u;

namespace {

enum { SITE_ATTN = 0, SITE_HID1 = 1, SITE_HID2 = 2 };

b200u_dropout_t site(const b200u_layer_params_t* p, int which, float prob) {
    b200u_dropout_t d;
    d.seed_ptr = p->seed;
    d.stream = p->stream_base + (uint32_t)which;
    d.p = prob;
    return d;
}

struct LnArgs {
    const float* gamma; const float* beta; float* mean; float* rstd; float eps;
};

int gemm(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn,
         int epi, void* C, int ldc, void* C2, int ldc2, const float* bias, const void* R, int ldr,
         const b200u_dropout_t* drop, int impl, cudaStream_t stream, float* colsum = nullptr,
         const LnArgs* ln = nullptr) {
    b200u_gemm_t g;
    memset(&g, 0, sizeof(g));
    g.colsum = colsum;
    if (ln) {
        g.ln_gamma = ln->gamma; g.ln_beta = ln->beta; g.ln_mean = ln->mean; g.ln_rstd = ln->rstd;
        g.ln_eps = ln->eps;
    }
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.lda = lda; g.a_mn_major = a_mn;
    g.B = B; g.ldb = ldb; g.b_mn_major = b_mn;
    g.epilogue = epi;
    g.C = C; g.ldc = ldc; g.C2 = C2; g.ldc2 = ldc2;
    g.bias = bias; g.R = R; g.ldr = ldr;
    if (drop) g.drop = *drop;
    g.impl = impl;
    return b200u_gemm(&g, stream);
}

#define TRY(expr)            \
    do {                     \
        int _rc = (expr);    \
        if (_rc) return _rc; \
    } while (0)

}  // namespace

// ---- side stream for the weight-gradient GEMMs of the backward pass -------------------------------
// dW GEMMs (and the b1 bias-gradient column sum) are off the critical path: nothing in the layer's
// backward chain consumes them. They run on a library-owned side stream, forked from / joined to the
// caller's stream with events (capturable: the side stream joins the capture through the fork event),
// so their CTAs fill the SMs that the critical-path kernels leave idle -- partial second waves of the
// persistent GEMMs (252 tiles on 148 SMs), the 126-CTA N=768 GEMMs, the 192-CTA attention backward.
namespace {

struct SideCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork[4] = {nullptr, nullptr, nullptr, nullptr};  // main -> side: producer of each side job is done
    cudaEvent_t w2_done = nullptr;                               // side -> main: dz may be overwritten
    cudaEvent_t all_done = nullptr;                              // side -> main: join at the end of the layer
    bool ok = false;
};
// b200u_set_bwd_streams(): 1 = weight gradients on the side stream, 0 = single stream
// (the environment variable B200U_BWD_STREAMS=0 selects the single-stream order at load time)
// b200u_set_fused_layernorm(): 1 (default) = LayerNorm inside the attn-out / FFN2 GEMM epilogues
// (B200U_FUSE_LN=0 selects the separate launches at load time)
int g_fuse_ln = [] {
    const char* e = getenv("B200U_FUSE_LN");
    return (e && e[0] == '0') ? 0 : 1;
}();
int g_side_mode = [] {
    const char* e = getenv("B200U_BWD_STREAMS");
    return (e && e[0] == '0') ? 0 : 1;
}();

SideCtx* side_ctx() {
    // one context per device, shared by all host threads (PyTorch runs backward on its own thread;
    // calls for one device are serialised by the caller's stream order)
    static SideCtx ctx[16];
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    SideCtx& c = ctx[dev];
    if (!c.ok) {
        if (cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int i = 0; i < 4; ++i)
            if (cudaEventCreateWithFlags(&c.fork[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&c.w2_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&c.all_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        c.ok = true;
    }
    return &c;
}

}  // namespace

extern "C" int b200u_set_fused_layernorm(int on) {
    g_fuse_ln = on ? 1 : 0;
    return B200U_OK;
}

extern "C" int b200u_set_bwd_streams(int two_streams) {
    g_side_mode = two_streams ? 1 : 0;
    return B200U_OK;
}

extern "C" int b200u_bert_layer_fwd(const b200u_layer_params_t* p, const void* x0,
                                    const b200u_layer_saved_t* s, void* x2, b200u_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    B200U_CHECK_ARG(p && x0 && s && x2, "bert_layer_fwd: null pointer");
    const int M = p->B * p->L, H = p->H, I = p->I;
    if (M == 0) return B200U_OK;
    const b200u_dropout_t d_attn = site(p, SITE_ATTN, p->p_attn);
    const b200u_dropout_t d_h1 = site(p, SITE_HID1, p->p_hidden);
    const b200u_dropout_t d_h2 = site(p, SITE_HID2, p->p_hidden);
    if (g_side_mode) side_ctx();  // create the backward's side stream / events outside any later capture

    TRY(gemm(M, 3 * H, H, x0, H, 0, p->Wqkv, H, 0, B200U_EPI_STORE, s->qkv, 3 * H, nullptr, 0, p->bqkv,
             nullptr, 0, nullptr, p->gemm_impl, st));
    TRY(b200u_attention_fwd(s->qkv, p->mask, s->ctx, s->lse, p->B, p->L, p->heads, H, &d_attn, st));
    const bool fuse_ln = g_fuse_ln && (H % 128 == 0) && (H / 128 <= 8);
    if (fuse_ln) {
        const LnArgs ln1 = {p->ln1_g, p->ln1_b, s->mean1, s->rstd1, p->eps};
        TRY(gemm(M, H, H, s->ctx, H, 0, p->Wo, H, 0, B200U_EPI_BIAS_DROP_RES_LN, s->y1, H, s->x1, H, p->bo, x0, H,
                 &d_h1, p->gemm_impl, st, nullptr, &ln1));
    } else {
        TRY(gemm(M, H, H, s->ctx, H, 0, p->Wo, H, 0, B200U_EPI_BIAS_DROP_RES, s->y1, H, nullptr, 0, p->bo, x0,
                 H, &d_h1, p->gemm_impl, st));
        TRY(b200u_layernorm_fwd(s->y1, B200U_BF16, p->ln1_g, p->ln1_b, s->x1, B200U_BF16, s->mean1, s->rstd1,
                                M, H, p->eps, nullptr, st));
    }
    // saved->u holds gelu'(u) (what the backward needs), saved->g = gelu(u)
    TRY(gemm(M, I, H, s->x1, H, 0, p->W1, H, 0, B200U_EPI_BIAS_GELU_DG, s->u, I, s->g, I, p->b1, nullptr, 0,
             nullptr, p->gemm_impl, st));
    if (fuse_ln) {
        const LnArgs ln2 = {p->ln2_g, p->ln2_b, s->mean2, s->rstd2, p->eps};
        TRY(gemm(M, H, I, s->g, I, 0, p->W2, I, 0, B200U_EPI_BIAS_DROP_RES_LN, s->y2, H, x2, H, p->b2, s->x1, H,
                 &d_h2, p->gemm_impl, st, nullptr, &ln2));
    } else {
        TRY(gemm(M, H, I, s->g, I, 0, p->W2, I, 0, B200U_EPI_BIAS_DROP_RES, s->y2, H, nullptr, 0, p->b2, s->x1,
                 H, &d_h2, p->gemm_impl, st));
        TRY(b200u_layernorm_fwd(s->y2, B200U_BF16, p->ln2_g, p->ln2_b, x2, B200U_BF16, s->mean2, s->rstd2, M,
                                H, p->eps, nullptr, st));
    }
    return B200U_OK;
}

extern "C" int b200u_bert_layer_bwd(const b200u_layer_params_t* p, const void* x0,
                                    const b200u_layer_saved_t* s, const void* dx2,
                                    const b200u_layer_grads_t* g, const b200u_layer_scratch_t* w,
                                    void* dx0, b200u_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    B200U_CHECK_ARG(p && x0 && s && dx2 && g && w && dx0, "bert_layer_bwd: null pointer");
    const int M = p->B * p->L, H = p->H, I = p->I;
    if (M == 0) return B200U_OK;
    const b200u_dropout_t d_attn = site(p, SITE_ATTN, p->p_attn);
    const b200u_dropout_t d_h1 = site(p, SITE_HID1, p->p_hidden);
    const b200u_dropout_t d_h2 = site(p, SITE_HID2, p->p_hidden);
    const int impl = p->gemm_impl;

    SideCtx* sc = g_side_mode ? side_ctx() : nullptr;
    cudaStream_t sd = sc ? sc->stream : st;  // where the weight-gradient work goes
    // fork k: the side stream may start job k once everything enqueued on `st` so far is done
#define FORK(k)                                                     \
    do {                                                            \
        if (sc) {                                                   \
            B200U_CHECK_CUDA(cudaEventRecord(sc->fork[k], st));     \
            B200U_CHECK_CUDA(cudaStreamWaitEvent(sd, sc->fork[k], 0)); \
        }                                                           \
    } while (0)

    // LN2 backward: dres (= grad of x1 through the residual) and dz2 (= grad of the FFN2 output)
    void* dz2 = p->p_hidden > 0.f ? w->dz : w->dres;
    TRY(b200u_layernorm_bwd(dx2, s->y2, B200U_BF16, s->mean2, s->rstd2, p->ln2_g, w->dres,
                            p->p_hidden > 0.f ? w->dz : nullptr, g->dln2_g, g->dln2_b, g->db2, M, H, &d_h2,
                            0, st));
    // FFN2: dW2[H,I] += dz2ᵀ·g (side) ; du = (dz2·W2) * gelu'(u) with gelu'(u) saved by the forward, and
    // db1 += colsum(du) folded into the same epilogue
    FORK(0);
    TRY(gemm(H, I, M, dz2, H, 1, s->g, I, 1, B200U_EPI_ATOMIC_F32, g->dW2, I, nullptr, 0, nullptr, nullptr,
             0, nullptr, impl, sd));
    if (sc) B200U_CHECK_CUDA(cudaEventRecord(sc->w2_done, sd));
    TRY(gemm(M, I, H, dz2, H, 0, p->W2, I, 1, B200U_EPI_MUL, w->du, I, nullptr, 0, nullptr, s->u, I,
             nullptr, impl, st, g->db1));
    // FFN1: dW1[I,H] += duᵀ·x1 (side) ; dx1 = du·W1 + dres
    FORK(1);
    TRY(gemm(I, H, M, w->du, I, 1, s->x1, H, 1, B200U_EPI_ATOMIC_F32, g->dW1, H, nullptr, 0, nullptr,
             nullptr, 0, nullptr, impl, sd));
    TRY(gemm(M, H, I, w->du, I, 0, p->W1, H, 1, B200U_EPI_ADD, w->dx1, H, nullptr, 0, nullptr, w->dres, H,
             nullptr, impl, st));
    // LN1 backward (rewrites dres / dz: the FFN2 weight gradient must have consumed dz2 first)
    if (sc) B200U_CHECK_CUDA(cudaStreamWaitEvent(st, sc->w2_done, 0));
    void* dz1 = p->p_hidden > 0.f ? w->dz : w->dres;
    TRY(b200u_layernorm_bwd(w->dx1, s->y1, B200U_BF16, s->mean1, s->rstd1, p->ln1_g, w->dres,
                            p->p_hidden > 0.f ? w->dz : nullptr, g->dln1_g, g->dln1_b, g->dbo, M, H, &d_h1,
                            0, st));
    // attention output projection: dWo[H,H] += dz1ᵀ·ctx (side) ; dctx = dz1·Wo
    FORK(2);
    TRY(gemm(H, H, M, dz1, H, 1, s->ctx, H, 1, B200U_EPI_ATOMIC_F32, g->dWo, H, nullptr, 0, nullptr,
             nullptr, 0, nullptr, impl, sd));
    TRY(gemm(M, H, H, dz1, H, 0, p->Wo, H, 1, B200U_EPI_STORE, w->dctx, H, nullptr, 0, nullptr, nullptr, 0,
             nullptr, impl, st));
    // (the QKV bias gradient, column sums of dqkv, is accumulated inside the attention backward)
    TRY(b200u_attention_bwd(s->qkv, p->mask, s->ctx, w->dctx, s->lse, w->dqkv, w->attn, g->dbqkv, p->B, p->L,
                            p->heads, H, &d_attn, st));
    // QKV projection: dWqkv[3H,H] += dqkvᵀ·x0 (side) ; dx0 = dqkv·Wqkv + dres
    FORK(3);
    TRY(gemm(3 * H, H, M, w->dqkv, 3 * H, 1, x0, H, 1, B200U_EPI_ATOMIC_F32, g->dWqkv, H, nullptr, 0,
             nullptr, nullptr, 0, nullptr, impl, sd));
    TRY(gemm(M, H, 3 * H, w->dqkv, 3 * H, 0, p->Wqkv, H, 1, B200U_EPI_ADD, dx0, H, nullptr, 0, nullptr,
             w->dres, H, nullptr, impl, st));
    if (sc) {
        // join: the scratch buffers the side jobs read (dz, du, dqkv) are rewritten by the next layer
        B200U_CHECK_CUDA(cudaEventRecord(sc->all_done, sd));
        B200U_CHECK_CUDA(cudaStreamWaitEvent(st, sc->all_done, 0));
    }
#undef FORK
    return B200U_OK;
}
