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

enum { B200U_BF16 = 0, B200U_F32 = 1 };

const char* b200u_last_error_string(void);
/* Input-error word of the current device. The index-consuming kernels (embedding lookups and their scatter
 * backward, the gather_index concat) never read or write out of bounds: an index outside its table makes them
 * OR one of these bits into a 4-byte device word (the only memory the library owns) and use a safe substitute
 * (row 0 / skip the row). b200u_input_errors synchronises the device, copies the word to *bits and optionally
 * clears it; the reference raises IndexError / a device assert in the same situations (torch nn.Embedding,
 * torch.gather at model/model.py:329-333). */
enum {
    B200U_ERR_WORD_ID = 1,
    B200U_ERR_POS_ID = 2,
    B200U_ERR_TYPE_ID = 4,
    B200U_ERR_GATHER_INDEX = 8,
    B200U_ERR_SCATTER_ID = 16
};
int b200u_input_errors(unsigned* bits /* host */, int reset);
int b200u_version(void);
/* Compiled SASS arch (100 for sm_100a) and SM count of the current device. */
int b200u_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Number of kernels this library has launched in the process so far (bench.py gpu_launches). */
long long b200u_launch_count(void);
/* Programmatic dependent launch (default on): every kernel of the library is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization and calls griddepcontrol.wait before its
 * first global-memory access, so its launch + prologue overlap the tail of the preceding kernel
 * on the stream (also inside captured CUDA graphs). 0 turns the attribute off. */
int b200u_set_pdl(int on);
/* Size persistent grids (GEMM tile loops, row kernels) for at most n SMs (0 = all). The data-parallel
 * step sets this while gradient all-reduces run beside the backward pass: NCCL's CTAs own their SMs,
 * and a 148-CTA persistent grid that cannot be fully resident would run as two waves. */
int b200u_set_sm_limit(int n);
/* Backward pass of b200u_bert_layer_bwd: 1 (default) = weight-gradient GEMMs run on a library-owned
 * side stream forked from / joined to the caller's stream with events (capturable), so they fill the
 * SMs the critical-path kernels leave idle; 0 = everything on the caller's stream. */
int b200u_set_bwd_streams(int two_streams);
/* b200u_bert_layer_fwd: 1 (default) = residual + LayerNorm inside the attn-out / FFN2 GEMM epilogues
 * (EPI_BIAS_DROP_RES_LN, 5 launches per layer); 0 = LayerNorm as its own launch (7 launches). */
int b200u_set_fused_layernorm(int on);
/* Event-time every tcgen05 GEMM launch (bench.py roofline leg). enable(n) arms up to n records,
 * collect() synchronises and returns summed duration / algorithmic FLOPs (2MNK) / record count.
 * Must not be armed during CUDA-graph capture. */
int b200u_prof_enable(int max_records);
int b200u_prof_collect(double* total_ms, double* total_flops, int* count);

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
    B200U_EPI_BIAS_GELU_DG = 7,  /* u = acc + bias ; C(bf16) = gelu_erf'(u) ; C2(bf16) = gelu_erf(u):
                                    the FFN1 forward saves the DERIVATIVE instead of the pre-activation,
                                    so the backward epilogue is a plain multiply (EPI_MUL)           */
    B200U_EPI_MUL = 8,           /* C(bf16) = acc * R ; optional colsum[n] += sum_m C(m,n) (bias grad) */
    B200U_EPI_BIAS_DROP_RES_LN = 9, /* y = dropout(acc + bias) + R -> C(bf16) ; C2(bf16) = LayerNorm(y)
                                    over the full row (N = 128 * cluster size <= 1024): the CTAs that own
                                    the N/128 column tiles of one row block form a thread-block cluster and
                                    exchange row statistics through distributed shared memory;
                                    ln_mean / ln_rstd (f32 [M]) are saved for the backward
                                    (model/layer.py:111-115,152-156)                                  */
    B200U_EPI_CE_STATS = 10,     /* vocabulary GEMM fused with cross entropy, forward (BertLMPredictionHead decoder +
                                    F.cross_entropy, model/layer.py:204-221, model/pretrain.py:97-98): the logits
                                    acc + bias are never written; per (row, 128-column tile) the epilogue emits
                                    ce_partial = (max, sum exp(x - max)) and ce_tlogit[row] = logit at ce_target[row].
                                    K-major A and B only; b200u_ce_finish turns the partials into lse and loss   */
    B200U_EPI_CE_GRAD = 11,      /* backward of the same: the logits are RECOMPUTED and
                                    C(bf16)[m,n] = (exp(acc + bias - ce_lse[m]) - [n == ce_target[m]]) * ce_scale[m],
                                    the operand of the decoder's dgrad / wgrad GEMMs                            */
    B200U_EPI_COUNT = 12
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
    int cluster;            /* 0 = auto, 1 = no cluster, 2 = CTA pairs with TMA-multicast B tiles */
    float* colsum;          /* EPI_MUL: f32 [N] or NULL, += column sums of the bf16 output */
    const float* ln_gamma;  /* EPI_BIAS_DROP_RES_LN: f32 [N] LayerNorm weight / bias, eps, saved statistics */
    const float* ln_beta;
    float* ln_mean;         /* f32 [M] or NULL */
    float* ln_rstd;         /* f32 [M] or NULL */
    float ln_eps;
    const long long* ce_target; /* EPI_CE_STATS / EPI_CE_GRAD: i64 [M] target column of each row */
    float* ce_partial;          /* EPI_CE_STATS: f32 [M][ceil(N/128)][2] (max, sum exp) per row and column tile */
    float* ce_tlogit;           /* EPI_CE_STATS: f32 [M] logit at the target */
    const float* ce_lse;        /* EPI_CE_GRAD: f32 [M] log-sum-exp of each row (from b200u_ce_finish) */
    const float* ce_scale;      /* EPI_CE_GRAD: f32 [M] upstream gradient of each row's loss */
} b200u_gemm_t;

int b200u_gemm(const b200u_gemm_t* g, b200u_stream_t stream);
/* Second half of EPI_CE_STATS: lse[m] = log sum_n exp(logit[m,n]) from the per-tile partials (n_tiles =
 * ceil(N/128) of the GEMM) and loss[m] = lse[m] - tlogit[m] (F.cross_entropy(..., reduction='none')). */
int b200u_ce_finish(const float* partial, const float* tlogit, float* lse, float* loss, int M, int n_tiles,
                    b200u_stream_t stream);
/* Bring-up aid: non-NULL -> every tcgen05 GEMM CTA writes 8 clock64() phase stamps (int64) to
 * device_ptr[cta*8 ..] (entry, setup done, TMA issued, first operands landed, MMAs issued,
 * accumulator ready, epilogue done, exit). NULL disables. */
int b200u_gemm_debug_stamps(long long* device_ptr);

/* ------------------------------------------------------------------------------------------
 * K5  LayerNorm (Apex FusedLayerNorm(H, eps=1e-12) replacement: model/model.py:229,252,253,258;
 * model/layer.py:108,149). fwd: y = (x-mean)*rstd*gamma+beta, optional dropout on y, saves
 * mean/rstd [M] for backward. bwd: dx (bf16), optionally dz = dropout_mask(dx) and
 * dbias += colsum(dz) for the `LN(dropout(dense(h)) + residual)` pattern
 * (model/layer.py:111-115,152-156); dgamma/dbeta/dbias are ACCUMULATED (+=) in f32.
 * drop_on_input != 0: the module is dropout(LN(x)) (embeddings), so dy is masked on load. */
int b200u_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y,
                        int y_dtype, float* mean, float* rstd, int M, int H, float eps,
                        const b200u_dropout_t* drop, b200u_stream_t stream);
int b200u_layernorm_bwd(const void* dy, const void* x, int x_dtype, const float* mean,
                        const float* rstd, const float* gamma, void* dx, void* dz, float* dgamma,
                        float* dbeta, float* dbias, int M, int H, const b200u_dropout_t* drop,
                        int drop_on_input, b200u_stream_t stream);

/* out[n] += sum_m x[m,n] (bf16 in, f32 accumulate): nn.Linear bias gradients. */
int b200u_colsum_accum(const void* x, int ldx, float* out, int M, int N, b200u_stream_t stream);
/* out = dy * gelu_erf'(u), bf16, n % 8 == 0 (backward of dense+GELU head transforms,
 * model/layer.py:196-200, model/pretrain.py:23-25). */
int b200u_dgelu_mul(const void* dy, const void* u, void* out, size_t n, b200u_stream_t stream);
/* y(bf16)[i] = x(f32)[i], n % 8 == 0: img_feat / weight shadow casts. */
int b200u_cast_f32_to_bf16(const float* x, void* y, size_t n, b200u_stream_t stream);
/* dst(bf16)[i] = dst[i] + sum_k srcs[k][i] (fp32 accumulation, sources in the order given), n % 8 == 0,
 * nsrc <= B200U_MAX_PEERS: the local reduce step of the data-parallel gradient exchange, where every rank
 * sums its own slice of a bucket with the copies its peers pushed over NVLink with the copy engines
 * (replaces the reduce half of the NCCL ring all-reduce behind torch DDP-style averaging,
 * train_template.py:95-103 has no DP in the reference). */
#define B200U_MAX_PEERS 15
int b200u_slice_sum_bf16(void* dst, const void* const* srcs, int nsrc, size_t n, b200u_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K1  gather_index concat (model/model.py:329-333): out[b,j,:] = cat(txt,img)[b, gather_index[b,j], :]
 * as 16-byte row copies, bit-exact including padded positions; txt [B,T,H], img [B,R,H],
 * gather_index int64 [B,L], out [B,L,H], all bf16. bwd = deterministic scatter-add. */
int b200u_gather_rows(const void* txt, const void* img, const long long* gather_index, void* out,
                      int B, int T, int R, int L, int H, b200u_stream_t stream);
int b200u_gather_rows_bwd(const void* dout, const long long* gather_index, void* dtxt, void* dimg,
                          int B, int T, int R, int L, int H, b200u_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K0  UniterTextEmbeddings.forward (model/model.py:232-245): LN(word[ids]+pos[pos_ids]+type[tids])
 * then dropout. position_ids may be [B,T] (pos_batch_stride = T) or [1,T] (stride 0);
 * type_ids NULL -> zeros. sum_out (f32 [B*T,H], pre-LN sum) and mean/rstd feed the backward.
 * *_rows = row counts of the three tables: an id outside [0, rows) (nn.Embedding raises IndexError) sets
 * B200U_ERR_* in the device's input-error word (b200u_input_errors) and row 0 is read instead. */
int b200u_txt_embed_fwd(const long long* input_ids, const long long* position_ids,
                        int pos_batch_stride, const long long* type_ids, const float* word,
                        const float* pos, const float* type, const float* gamma, const float* beta,
                        void* out, float* sum_out, float* mean, float* rstd, int B, int T, int H,
                        int vocab_rows, int pos_rows, int type_rows, float eps,
                        const b200u_dropout_t* drop, b200u_stream_t stream);
/* K2  UniterImageEmbeddings.forward after img_linear (model/model.py:267-271):
 * LN(LN_img(a) + LN_pos(pos7·Wposᵀ+bpos) + type[type_ids]) then dropout. a = img_linear output
 * (f32 [n,H], from the GEMM). type_ids NULL -> ones (model/model.py:313-314).
 * p_out/s_out (f32 [n,H]) and stats_out (f32 [6,n]: mean/rstd of the three LNs) feed the backward. */
int b200u_img_embed_fwd(const float* a, const float* pos7, const float* Wpos, const float* bpos,
                        const long long* type_ids, const float* type, const float* g_img,
                        const float* b_img, const float* g_pos, const float* b_pos, const float* g,
                        const float* b, void* out, float* p_out, float* s_out, float* stats_out,
                        int n, int H, int type_rows, float eps, const b200u_dropout_t* drop,
                        b200u_stream_t stream);
/* nn.Embedding backward: table_grad[id(r), :] += d[r, :] (bf16 rows, f32 atomics). ids NULL ->
 * every row uses const_id; ids are indexed [b*ids_batch_stride + t] with r = b*T + t (n = B*T). Rows whose id
 * is outside [0, rows) are skipped and flagged (B200U_ERR_SCATTER_ID): nothing is ever added outside the table. */
int b200u_embedding_scatter_add(const void* d, const long long* ids, int ids_batch_stride, int T,
                                long long const_id, float* table_grad, int n, int H,
                                long long padding_idx, long long rows, b200u_stream_t stream);
/* Deterministic (atomic-free) form for rows sorted by id: ids_sorted ascending, perm[p] = source row of
 * sorted entry p. One read-modify-write per distinct id, rows of a run added in sorted order, so equal
 * inputs give bit-identical tables (data-parallel exchange of the touched word-embedding rows).
 * sumsq (f64 [sumsq_slots] or NULL): slot (block % sumsq_slots) += sum of squares of the rows written; the slots'
 * total is the table's share of the gradient norm when the table was zero before the call (the fused optimizer
 * clears it every step). Several slots keep thousands of warps from serialising on one L2 address. */
int b200u_embedding_segment_add(const void* d, const long long* ids_sorted, const long long* perm,
                                float* table_grad, int n, int H, long long padding_idx, long long rows,
                                double* sumsq, int sumsq_slots, b200u_stream_t stream);
/* pos_linear weight gradient: dW[h,c] += sum_r dp[r,h] * pos7[r,c]. */
int b200u_pos_linear_wgrad(const void* dp, const float* pos7, float* dW, int n, int H,
                           b200u_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4  BertSelfAttention core (model/layer.py:80-100) on the fused qkv activation [B*L, 3H]:
 * softmax(Q·Kᵀ/8 + mask) -> dropout -> ·V -> ctx [B*L, H]. mask f32 [B,L] is the ADDITIVE
 * mask UniterModel.forward builds, (1-attention_mask)*-10000 (model/model.py:342-345), one row
 * per sample (the [B,1,1,L] broadcast is never materialised). head_dim must be 64, L<=256.
 * lse f32 [B,heads,L] is saved for the backward, which writes dqkv [B*L, 3H]. */
int b200u_attention_fwd(const void* qkv, const float* mask, void* ctx, float* lse, int B, int L,
                        int num_heads, int H, const b200u_dropout_t* drop, b200u_stream_t stream);
/* 1 (default) = tcgen05 / TMEM / TMA attention kernels (one CTA per sample x head x 128-query tile, thread =
 * query row softmax); 0 = the warp-level mma.sync kernels (bring-up reference, differential tests). */
int b200u_set_attention_impl(int tcgen05);
/* scratch: caller-owned device buffer of b200u_attention_bwd_scratch_bytes(B, L, heads) bytes
 * (bf16 probabilities and score gradients handed from the dQ launch to the dK/dV launch; only
 * touched when L > 176, shorter sequences run the fused single-kernel backward).
 * dbias_qkv (f32 [3H], may be NULL): += column sums of the bf16 dqkv, i.e. the bias gradient of the
 * fused QKV projection (model/layer.py:64-66), accumulated inside the same launch. */
size_t b200u_attention_bwd_scratch_bytes(int B, int L, int num_heads);
int b200u_attention_bwd(const void* qkv, const float* mask, const void* ctx, const void* dctx,
                        const float* lse, void* dqkv, void* scratch, float* dbias_qkv, int B, int L,
                        int num_heads, int H, const b200u_dropout_t* drop, b200u_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * BertLayer.forward / its backward as one call each (model/layer.py:159-170). Weight matrices
 * are the bf16 shadows ([out,in] like nn.Linear.weight; Wqkv = cat(query,key,value).weight),
 * biases / LayerNorm params f32. `saved` tensors are written by fwd and read by bwd; `grads`
 * are f32 and ACCUMULATED; `scratch` is reusable across layers. dx0 may alias dx2. */
typedef struct {
    int B, L, H, I, heads;
    float eps;
    const void* Wqkv; const float* bqkv; /* [3H,H], [3H] */
    const void* Wo;   const float* bo;   /* [H,H],  [H]  */
    const float* ln1_g; const float* ln1_b;
    const void* W1;   const float* b1;   /* [I,H],  [I]  */
    const void* W2;   const float* b2;   /* [H,I],  [H]  */
    const float* ln2_g; const float* ln2_b;
    const float* mask;                   /* f32 [B,L] additive mask (0 / -10000) */
    float p_attn, p_hidden;              /* dropout probabilities (0 in eval mode) */
    const unsigned long long* seed;      /* device seed, required when any p > 0 */
    uint32_t stream_base;                /* RNG stream id of this layer's first dropout site */
    int gemm_impl;                       /* 0 = tcgen05, 1 = SIMT debug */
} b200u_layer_params_t;
typedef struct {
    void* qkv;  /* bf16 [M,3H] */
    void* ctx;  /* bf16 [M,H]  */
    float* lse; /* f32 [B,heads,L] */
    void* y1;   /* bf16 [M,H] pre-LN1 */
    float* mean1; float* rstd1;
    void* x1;   /* bf16 [M,H] LN1 out */
    void* u;    /* bf16 [M,I] gelu'(pre-activation): the derivative, not the pre-activation, is saved */
    void* g;    /* bf16 [M,I] post-GELU */
    void* y2;   /* bf16 [M,H] pre-LN2 */
    float* mean2; float* rstd2;
} b200u_layer_saved_t;
typedef struct {
    float* dWqkv; float* dbqkv; float* dWo; float* dbo; float* dln1_g; float* dln1_b;
    float* dW1; float* db1; float* dW2; float* db2; float* dln2_g; float* dln2_b;
} b200u_layer_grads_t;
typedef struct {
    void* dres; /* bf16 [M,H]  */
    void* dz;   /* bf16 [M,H]  */
    void* dx1;  /* bf16 [M,H]  */
    void* dctx; /* bf16 [M,H]  */
    void* du;   /* bf16 [M,I]  */
    void* dqkv; /* bf16 [M,3H] */
    void* attn; /* b200u_attention_bwd_scratch_bytes(B, L, heads) bytes */
} b200u_layer_scratch_t;
int b200u_bert_layer_fwd(const b200u_layer_params_t* p, const void* x0, const b200u_layer_saved_t* saved,
                         void* x2, b200u_stream_t stream);
int b200u_bert_layer_bwd(const b200u_layer_params_t* p, const void* x0, const b200u_layer_saved_t* saved,
                         const void* dx2, const b200u_layer_grads_t* grads,
                         const b200u_layer_scratch_t* scratch, void* dx0, b200u_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K7  BertPooler (model/layer.py:179-185), small heads (model/meme_uniter.py:20,
 * model/pretrain.py:62) and BCEWithLogitsLoss(pos_weight) (train_template.py:64-65,98-99).
 * h rows are bf16 at h + b*row_stride (first token of sample b); weights f32. Parameter grads
 * are accumulated (+=). */
int b200u_pooler_fwd(const void* h, long long row_stride, const float* W, const float* bias,
                     float* pooled, int B, int H, b200u_stream_t stream);
int b200u_pooler_bwd(const float* dpooled, const float* pooled, const void* h, long long row_stride,
                     const float* W, float* dW, float* db, void* dh, long long dh_row_stride, int B,
                     int H, b200u_stream_t stream);
int b200u_linear_small_fwd(const float* x, const float* W, const float* bias, float* out, int B,
                           int C, int K, b200u_stream_t stream);
int b200u_linear_small_bwd(const float* dout, const float* x, const float* W, float* dx, float* dW,
                           float* db, int B, int C, int K, b200u_stream_t stream);
/* loss = mean BCE-with-logits; dlogits = d(loss*grad_scale)/dlogits; probs = sigmoid(logits).
 * Any of loss/dlogits/probs may be NULL. */
int b200u_bce_logits(const float* logits, const float* labels, float pos_weight, float grad_scale,
                     float* loss, float* dlogits, float* probs, int B, b200u_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K8  optimal transport (model/ot.py), fp32. x [B,M,D] text, y [B,N,D] regions, pads uint8
 * (1 = padding) [B,M] / [B,N].
 *  cosine_cost : cost[B,M,N] = 1 - normalize(x)·normalize(y)ᵀ with joint padding zeroed
 *                (ot.py:11-21,72-75); xinv/yinv = 1/max(norm, eps) are kept for the backward.
 *  ipot        : T[B,N,M] after `iterations` IPOT steps, one launch (ot.py:35-66).
 *  ot_distance : dist[B] = trace(cost · T) (ot.py:24-32,84).
 *  cosine_cost_bwd : d dist / d x, d y with T detached (ot.py:82-84). */
int b200u_cosine_cost(const float* x, const float* y, const unsigned char* x_pad,
                      const unsigned char* y_pad, float* cost, float* xinv, float* yinv, int B, int M,
                      int N, int D, float eps, b200u_stream_t stream);
int b200u_ipot(const float* cost, const unsigned char* x_pad, const unsigned char* y_pad, float* T,
               int B, int M, int N, float beta, int iterations, int k, b200u_stream_t stream);
int b200u_ot_distance(const float* cost, const float* T, float* dist, int B, int M, int N,
                      b200u_stream_t stream);
int b200u_cosine_cost_bwd(const float* x, const float* y, const float* xinv, const float* yinv,
                          const unsigned char* x_pad, const unsigned char* y_pad, const float* T,
                          const float* ddist, float* dx, float* dy, int B, int M, int N, int D,
                          b200u_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K10 fused optimizer step over flat f32 buffers (train_template.py:89-107,
 * utils/optim_utils.py:16-46): grad averaging + clip_grad_norm_ + Adam with L2 weight decay,
 * refreshing the bf16 weight shadow. lr / step / coef live in device memory (graph replay). */
int b200u_counter_add(unsigned long long* counter, unsigned long long inc, b200u_stream_t stream);
/* g16 / [lo16, hi16): optional bf16 gradient source for that element range (the data-parallel step
 * all-reduces the encoder-layer buckets in bf16 and leaves them there); NULL = everything from g.
 * run_wd[r] < 0 marks a run the optimizer skips (no gradient / frozen): only its gradient is zeroed. */
int b200u_grad_sumsq(const float* g, size_t n, double* sumsq, const void* g16, size_t lo16, size_t hi16,
                     b200u_stream_t stream);
int b200u_clip_coef(const double* sumsq, float pre_scale, float max_norm, float* coef,
                    float* norm_out, b200u_stream_t stream);
int b200u_adam_step(float* p, float* g, float* m, float* v, void* shadow_bf16, size_t n,
                    const long long* run_start, const float* run_wd, const int* chunk_run,
                    int num_runs, const float* coef, const float* lr,
                    const unsigned long long* step, float beta1, float beta2, float eps,
                    int zero_grad, const void* g16, size_t lo16, size_t hi16, b200u_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200U_H_ */
