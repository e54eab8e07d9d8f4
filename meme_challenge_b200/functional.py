"""torch.autograd.Functions over the b200u C-ABI.

Forward and backward of every Function are hand-written CUDA (no eager PyTorch arithmetic on the
hot path). Parameter gradients are ACCUMULATED in place into `param.grad` (views of the flat
gradient buffer when the owning module is flattened, see flat.py) and the Functions return None
for parameter inputs — the same end state `loss.backward()` leaves in the reference
(train_template.py:101-109), without per-tensor autograd accumulation kernels.
"""
import ctypes as C

import torch

from . import _lib, ops
from ._lib import (EPI_ATOMIC_F32, EPI_STORE_F32, DropoutT)

P = _lib.ptr


def grad_buf(p):
    """fp32 accumulation target for parameter p (created zeroed if absent). Records that p receives a
    gradient (the fused optimizer leaves parameters that never do untouched, like torch.optim.Adam skips
    tensors whose .grad is None)."""
    from .flat import store_of
    st = store_of(p)
    if st is not None:
        st.touched.add(id(p))
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


class Runtime(object):
    """Per-forward context shared by the Functions of one UniterModel call."""

    def __init__(self, store, training, seed, p_hidden, p_attn, gemm_impl=0, eps=1e-12):
        self.store = store
        self.training = training
        self.seed = seed              # int64 device tensor [1] or None
        self.p_hidden = p_hidden if training else 0.0
        self.p_attn = p_attn if training else 0.0
        self.gemm_impl = gemm_impl
        self.eps = eps
        self.layer_cb = None
        self.sparse_word_cb = None   # data parallel: takes (d_rows bf16 [n,H], ids [n], padding_idx) instead of the dense scatter

    def drop(self, stream_id, p):
        return _lib.dropout_t(self.seed, stream_id, p)


# ------------------------------------------------------------------------------------------------
# C structs of include/b200u.h (BertLayer composite)
# ------------------------------------------------------------------------------------------------
class LayerParamsT(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("L", C.c_int), ("H", C.c_int), ("I", C.c_int), ("heads", C.c_int),
        ("eps", C.c_float),
        ("Wqkv", C.c_void_p), ("bqkv", C.c_void_p),
        ("Wo", C.c_void_p), ("bo", C.c_void_p),
        ("ln1_g", C.c_void_p), ("ln1_b", C.c_void_p),
        ("W1", C.c_void_p), ("b1", C.c_void_p),
        ("W2", C.c_void_p), ("b2", C.c_void_p),
        ("ln2_g", C.c_void_p), ("ln2_b", C.c_void_p),
        ("mask", C.c_void_p),
        ("p_attn", C.c_float), ("p_hidden", C.c_float),
        ("seed", C.c_void_p),
        ("stream_base", C.c_uint32),
        ("gemm_impl", C.c_int),
    ]


class LayerSavedT(C.Structure):
    _fields_ = [("qkv", C.c_void_p), ("ctx", C.c_void_p), ("lse", C.c_void_p), ("y1", C.c_void_p),
                ("mean1", C.c_void_p), ("rstd1", C.c_void_p), ("x1", C.c_void_p), ("u", C.c_void_p),
                ("g", C.c_void_p), ("y2", C.c_void_p), ("mean2", C.c_void_p), ("rstd2", C.c_void_p)]


class LayerGradsT(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dWqkv", "dbqkv", "dWo", "dbo", "dln1_g", "dln1_b", "dW1",
                                          "db1", "dW2", "db2", "dln2_g", "dln2_b")]


class LayerScratchT(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dres", "dz", "dx1", "dctx", "du", "dqkv", "attn")]


def _layer_params(layer, rt, B, L, mask_add, layer_idx):
    st = rt.store
    att, out_att = layer.attention.self, layer.attention.output
    H = att.all_head_size
    I = layer.intermediate.dense.out_features
    p = LayerParamsT()
    p.B, p.L, p.H, p.I, p.heads = B, L, H, I, att.num_attention_heads
    p.eps = out_att.LayerNorm.eps
    p.Wqkv = st.fused(st.shadow, att.query.weight, 3 * H).data_ptr()
    p.bqkv = st.fused(st.flat, att.query.bias, 3 * H).data_ptr()
    p.Wo = st.w16(out_att.dense.weight).data_ptr()
    p.bo = out_att.dense.bias.data_ptr()
    p.ln1_g = out_att.LayerNorm.weight.data_ptr()
    p.ln1_b = out_att.LayerNorm.bias.data_ptr()
    p.W1 = st.w16(layer.intermediate.dense.weight).data_ptr()
    p.b1 = layer.intermediate.dense.bias.data_ptr()
    p.W2 = st.w16(layer.output.dense.weight).data_ptr()
    p.b2 = layer.output.dense.bias.data_ptr()
    p.ln2_g = layer.output.LayerNorm.weight.data_ptr()
    p.ln2_b = layer.output.LayerNorm.bias.data_ptr()
    p.mask = mask_add.data_ptr()
    p.p_attn, p.p_hidden = rt.p_attn, rt.p_hidden
    p.seed = rt.seed.data_ptr() if rt.seed is not None else None
    p.stream_base = 16 + 4 * layer_idx
    p.gemm_impl = rt.gemm_impl
    return p


def _layer_grads(layer, st):
    att, out_att = layer.attention.self, layer.attention.output
    H = att.all_head_size
    g = LayerGradsT()
    g.dWqkv = st.fused(st.grad, att.query.weight, 3 * H).data_ptr()
    g.dbqkv = st.fused(st.grad, att.query.bias, 3 * H).data_ptr()
    g.dWo = st.g32(out_att.dense.weight).data_ptr()
    g.dbo = st.g32(out_att.dense.bias).data_ptr()
    g.dln1_g = st.g32(out_att.LayerNorm.weight).data_ptr()
    g.dln1_b = st.g32(out_att.LayerNorm.bias).data_ptr()
    g.dW1 = st.g32(layer.intermediate.dense.weight).data_ptr()
    g.db1 = st.g32(layer.intermediate.dense.bias).data_ptr()
    g.dW2 = st.g32(layer.output.dense.weight).data_ptr()
    g.db2 = st.g32(layer.output.dense.bias).data_ptr()
    g.dln2_g = st.g32(layer.output.LayerNorm.weight).data_ptr()
    g.dln2_b = st.g32(layer.output.LayerNorm.bias).data_ptr()
    st.touched.update(id(q) for q in layer.parameters())
    return g


def _alloc_saved(M, H, I, B, heads, L, dev):
    bf = torch.empty(M * (3 * H + 4 * H + 2 * I), device=dev, dtype=torch.bfloat16)
    f32 = torch.empty(B * heads * L + 4 * M, device=dev, dtype=torch.float32)
    s = LayerSavedT()
    o = 0

    def take(n):
        nonlocal o
        v = bf[o:o + n]
        o += n
        return v
    t = {}
    t["qkv"] = take(M * 3 * H); t["ctx"] = take(M * H); t["y1"] = take(M * H); t["x1"] = take(M * H)
    t["u"] = take(M * I); t["g"] = take(M * I); t["y2"] = take(M * H)
    nl = B * heads * L
    t["lse"] = f32[0:nl]
    t["mean1"] = f32[nl:nl + M]; t["rstd1"] = f32[nl + M:nl + 2 * M]
    t["mean2"] = f32[nl + 2 * M:nl + 3 * M]; t["rstd2"] = f32[nl + 3 * M:nl + 4 * M]
    for k, v in t.items():
        setattr(s, k, v.data_ptr())
    return s, (bf, f32), t


_scratch_cache = {}


def _scratch(M, H, I, dev, B=0, L=0, heads=0):
    key = (M, H, I, B, L, heads, dev.index, torch.cuda.is_current_stream_capturing())
    hit = _scratch_cache.get(key)
    if hit is None:
        buf = torch.empty(M * (4 * H + I + 3 * H), device=dev, dtype=torch.bfloat16)
        w = LayerScratchT()
        o = 0
        for name, n in (("dres", M * H), ("dz", M * H), ("dx1", M * H), ("dctx", M * H), ("du", M * I),
                        ("dqkv", M * 3 * H)):
            setattr(w, name, buf[o:o + n].data_ptr())
            o += n
        attn = torch.empty(_lib.lib().b200u_attention_bwd_scratch_bytes(B, L, heads), device=dev, dtype=torch.uint8)
        w.attn = attn.data_ptr()
        hit = (w, buf, attn)
        if len(_scratch_cache) > 8:
            _scratch_cache.clear()
        _scratch_cache[key] = hit
    return hit[0]


class BertLayerFn(torch.autograd.Function):
    """model/layer.py:159-170 forward + backward through b200u_bert_layer_{fwd,bwd}."""

    @staticmethod
    def forward(ctx, x0, mask_add, anchor, layer, layer_idx, rt):
        B, L, H = x0.shape
        assert x0.dtype == torch.bfloat16 and x0.is_contiguous()
        I = layer.intermediate.dense.out_features
        M = B * L
        p = _layer_params(layer, rt, B, L, mask_add, layer_idx)
        saved, keep, _ = _alloc_saved(M, H, I, B, p.heads, L, x0.device)
        x2 = torch.empty_like(x0)
        _lib.check(_lib.lib().b200u_bert_layer_fwd(C.byref(p), P(x0), C.byref(saved), P(x2),
                                                   _lib.stream_ptr()), "b200u_bert_layer_fwd")
        ctx.save_for_backward(x0)
        ctx.p, ctx.saved, ctx.keep = p, saved, keep
        ctx.layer, ctx.rt, ctx.mask_add = layer, rt, mask_add
        return x2

    @staticmethod
    def backward(ctx, dx2):
        (x0,) = ctx.saved_tensors
        p, rt = ctx.p, ctx.rt
        if ctx.keep is None:
            # the saved activations (raw device pointers inside ctx.saved) are released after the first backward
            raise _lib.B200UError("BertLayer backward was already run for this forward (retain_graph / a second "
                                  "backward is not supported: the saved activations are freed after the first one)")
        dx2 = dx2.contiguous()
        g = _layer_grads(ctx.layer, rt.store)
        w = _scratch(p.B * p.L, p.H, p.I, x0.device, p.B, p.L, p.heads)
        dx0 = torch.empty_like(x0)
        _lib.check(_lib.lib().b200u_bert_layer_bwd(C.byref(p), P(x0), C.byref(ctx.saved), P(dx2),
                                                   C.byref(g), C.byref(w), P(dx0), _lib.stream_ptr()),
                   "b200u_bert_layer_bwd")
        ctx.keep = None
        return dx0, None, None, None, None, None


def bert_layer_infer(x0, mask_add, layer, layer_idx, rt, saved_cache):
    """No-grad forward: activations go to a reusable scratch instead of per-layer saved buffers."""
    B, L, H = x0.shape
    I = layer.intermediate.dense.out_features
    p = _layer_params(layer, rt, B, L, mask_add, layer_idx)
    key = (B, L, H, I, x0.device.index)
    if saved_cache.get("key") != key:
        saved_cache["key"] = key
        saved_cache["val"] = _alloc_saved(B * L, H, I, B, p.heads, L, x0.device)
    saved = saved_cache["val"][0]
    x2 = torch.empty_like(x0)
    _lib.check(_lib.lib().b200u_bert_layer_fwd(C.byref(p), P(x0), C.byref(saved), P(x2),
                                               _lib.stream_ptr()), "b200u_bert_layer_fwd")
    return x2


# ------------------------------------------------------------------------------------------------
def _ids64(t, name):
    """Index tensors reach the kernels as contiguous int64 (they read `const long long*`); integer tensors of
    another width are converted, anything else is rejected like nn.Embedding does."""
    if t.dtype == torch.int64:
        return t.contiguous()
    if t.dtype in (torch.int32, torch.int16, torch.int8, torch.uint8):
        return t.long().contiguous()
    raise _lib.B200UError("%s must be an integer tensor, got %s" % (name, t.dtype))


_ERR_BITS = ((1, "input_ids outside the word-embedding table"), (2, "position_ids outside the position table"),
             (4, "token type ids outside the token-type table"), (8, "gather_index outside [0, T + R)"),
             (16, "an embedding gradient row outside its table"))


def check_input_errors(reset=True):
    """Raise if any index-consuming kernel since the last check saw an out-of-range index (the kernels flag it and
    substitute a safe row instead of reading or writing out of bounds; the reference raises IndexError / a device
    assert). Synchronises the device: call it where the host waits anyway (after reading a loss, at evaluation end,
    before a checkpoint)."""
    bits = C.c_uint(0)
    _lib.check(_lib.lib().b200u_input_errors(C.byref(bits), int(bool(reset))), "b200u_input_errors")
    if bits.value:
        raise _lib.B200UError("invalid indices reached the device: " +
                              "; ".join(msg for bit, msg in _ERR_BITS if bits.value & bit))


class TxtEmbedFn(torch.autograd.Function):
    """UniterTextEmbeddings.forward (model/model.py:232-245)."""

    @staticmethod
    def forward(ctx, anchor, emb, input_ids, position_ids, token_type_ids, rt):
        B, T = input_ids.shape
        H = emb.word_embeddings.weight.shape[1]
        dev = input_ids.device
        # the kernels read 64-bit ids (nn.Embedding also accepts int32): convert anything else
        input_ids = _ids64(input_ids, "input_ids")
        position_ids = _ids64(position_ids, "position_ids")
        if position_ids.dim() == 1:
            position_ids = position_ids.unsqueeze(0)
        if position_ids.shape[0] not in (1, B) or position_ids.shape[1] != T:
            raise _lib.B200UError("position_ids must be [B,T] or [1,T]")
        pos_stride = T if position_ids.shape[0] == B else 0
        if token_type_ids is not None:
            token_type_ids = _ids64(token_type_ids, "token_type_ids")
        out = torch.empty(B, T, H, device=dev, dtype=torch.bfloat16)
        need = any(ctx.needs_input_grad)  # grad mode is off inside Function.forward
        sum_out = torch.empty(B * T, H, device=dev, dtype=torch.float32) if need else None
        mean = torch.empty(B * T, device=dev, dtype=torch.float32) if need else None
        rstd = torch.empty(B * T, device=dev, dtype=torch.float32) if need else None
        drop = rt.drop(1, rt.p_hidden)
        ops._call("b200u_txt_embed_fwd", P(input_ids), P(position_ids), pos_stride, P(token_type_ids),
                  P(emb.word_embeddings.weight), P(emb.position_embeddings.weight),
                  P(emb.token_type_embeddings.weight), P(emb.LayerNorm.weight), P(emb.LayerNorm.bias),
                  P(out), P(sum_out), P(mean), P(rstd), B, T, H, emb.word_embeddings.weight.shape[0],
                  emb.position_embeddings.weight.shape[0], emb.token_type_embeddings.weight.shape[0],
                  float(emb.LayerNorm.eps), C.byref(drop))
        ctx.emb, ctx.rt, ctx.drop = emb, rt, drop
        ctx.ids = (input_ids, position_ids, pos_stride, token_type_ids)
        ctx.stats = (sum_out, mean, rstd)
        return out

    @staticmethod
    def backward(ctx, dout):
        emb = ctx.emb
        input_ids, position_ids, pos_stride, token_type_ids = ctx.ids
        sum_out, mean, rstd = ctx.stats
        B, T = input_ids.shape
        H = sum_out.shape[1]
        dout = dout.contiguous().view(B * T, H)
        dx, _ = ops.layernorm_bwd(dout, sum_out, mean, rstd, emb.LayerNorm.weight,
                                  grad_buf(emb.LayerNorm.weight), grad_buf(emb.LayerNorm.bias),
                                  drop=ctx.drop, drop_on_input=True)
        pad = emb.word_embeddings.padding_idx
        if ctx.rt.sparse_word_cb is not None:
            # data-parallel last micro-batch: the (at most B*T) touched rows are exchanged between ranks
            # instead of all-reducing the dense [vocab, H] gradient after the backward pass
            grad_buf(emb.word_embeddings.weight)
            ctx.rt.sparse_word_cb(dx, input_ids.reshape(-1), -1 if pad is None else pad)
        else:
            ops._call("b200u_embedding_scatter_add", P(dx), P(input_ids), T, T, C.c_longlong(0),
                      P(grad_buf(emb.word_embeddings.weight)), B * T, H, C.c_longlong(-1 if pad is None else pad),
                      C.c_longlong(emb.word_embeddings.weight.shape[0]))
        ops._call("b200u_embedding_scatter_add", P(dx), P(position_ids), pos_stride, T, C.c_longlong(0),
                  P(grad_buf(emb.position_embeddings.weight)), B * T, H, C.c_longlong(-1),
                  C.c_longlong(emb.position_embeddings.weight.shape[0]))
        tg = grad_buf(emb.token_type_embeddings.weight)
        if token_type_ids is None:
            ops.colsum_accum(dx, tg[0])
        else:
            ops._call("b200u_embedding_scatter_add", P(dx), P(token_type_ids), T, T, C.c_longlong(0), P(tg),
                      B * T, H, C.c_longlong(-1), C.c_longlong(tg.shape[0]))
        return None, None, None, None, None, None


class ImgEmbedFn(torch.autograd.Function):
    """UniterImageEmbeddings.forward incl. the token-type lookup of _compute_img_embeddings
    (model/model.py:261-272, 311-319)."""

    @staticmethod
    def forward(ctx, anchor, iemb, type_table, img_feat, img_pos_feat, img_type_ids, img_masks, rt):
        B, R, D = img_feat.shape
        H = iemb.img_linear.out_features
        dev = img_feat.device
        n = B * R
        st = rt.store
        feat = img_feat
        if img_masks is not None:
            # model/model.py:262-265 (rare pretraining path: MRFR region masking)
            with torch.no_grad():
                iemb.mask_embedding.weight.data[0, :].fill_(0)
                feat = img_feat + iemb.mask_embedding.weight[img_masks.long()]
        feat = feat.contiguous().float()
        feat16 = torch.empty(n, D, device=dev, dtype=torch.bfloat16)
        ops.cast_f32_to_bf16(feat.view(-1), feat16.view(-1))
        w16 = st.w16(iemb.img_linear.weight)
        a = ops.gemm(feat16, w16, epilogue=EPI_STORE_F32, bias=iemb.img_linear.bias, impl=rt.gemm_impl)
        pos7 = img_pos_feat.contiguous().float().view(n, 7)
        if img_type_ids is not None:
            img_type_ids = _ids64(img_type_ids, "img_type_ids")
        out = torch.empty(B, R, H, device=dev, dtype=torch.bfloat16)
        need = any(ctx.needs_input_grad)  # grad mode is off inside Function.forward
        p_out = torch.empty(n, H, device=dev, dtype=torch.float32) if need else None
        s_out = torch.empty(n, H, device=dev, dtype=torch.float32) if need else None
        stats = torch.empty(6, n, device=dev, dtype=torch.float32) if need else None
        drop = rt.drop(2, rt.p_hidden)
        ops._call("b200u_img_embed_fwd", P(a), P(pos7), P(iemb.pos_linear.weight), P(iemb.pos_linear.bias),
                  P(img_type_ids), P(type_table), P(iemb.img_layer_norm.weight), P(iemb.img_layer_norm.bias),
                  P(iemb.pos_layer_norm.weight), P(iemb.pos_layer_norm.bias), P(iemb.LayerNorm.weight),
                  P(iemb.LayerNorm.bias), P(out), P(p_out), P(s_out), P(stats), n, H, type_table.shape[0],
                  float(iemb.LayerNorm.eps), C.byref(drop))
        ctx.iemb, ctx.rt, ctx.drop, ctx.type_table = iemb, rt, drop, type_table
        ctx.t = (feat16, a, pos7, p_out, s_out, stats, img_type_ids, img_masks)
        ctx.shape = (B, R, H, D)
        return out

    @staticmethod
    def backward(ctx, dout):
        iemb, rt = ctx.iemb, ctx.rt
        feat16, a, pos7, p_out, s_out, stats, img_type_ids, img_masks = ctx.t
        B, R, H, D = ctx.shape
        n = B * R
        dout = dout.contiguous().view(n, H)
        # LN (outer) backward, dropout applied to the incoming grad
        ds, _ = ops.layernorm_bwd(dout, s_out, stats[4], stats[5], iemb.LayerNorm.weight,
                                  grad_buf(iemb.LayerNorm.weight), grad_buf(iemb.LayerNorm.bias),
                                  drop=ctx.drop, drop_on_input=True)
        tg = grad_buf(ctx.type_table)
        if img_type_ids is None:
            ops.colsum_accum(ds, tg[1])
        else:
            ops._call("b200u_embedding_scatter_add", P(ds), P(img_type_ids), R, R, C.c_longlong(0), P(tg), n, H,
                      C.c_longlong(-1), C.c_longlong(tg.shape[0]))
        # LN_img backward -> da (grad of img_linear output) + its bias grad
        da, _ = ops.layernorm_bwd(ds, a, stats[0], stats[1], iemb.img_layer_norm.weight,
                                  grad_buf(iemb.img_layer_norm.weight), grad_buf(iemb.img_layer_norm.bias),
                                  dbias=grad_buf(iemb.img_linear.bias))
        # LN_pos backward -> dp (grad of pos_linear output) + its bias grad
        dp, _ = ops.layernorm_bwd(ds, p_out, stats[2], stats[3], iemb.pos_layer_norm.weight,
                                  grad_buf(iemb.pos_layer_norm.weight), grad_buf(iemb.pos_layer_norm.bias),
                                  dbias=grad_buf(iemb.pos_linear.bias))
        ops._call("b200u_pos_linear_wgrad", P(dp), P(pos7), P(grad_buf(iemb.pos_linear.weight)), n, H)
        # img_linear weight grad: dW[H,D] += daᵀ · feat
        ops.gemm(da, feat16, a_mn=True, b_mn=True, epilogue=EPI_ATOMIC_F32,
                 out=grad_buf(iemb.img_linear.weight), impl=rt.gemm_impl)
        if img_masks is not None:
            # d mask_embedding[1] = sum over masked regions of d img_feat = (da · W_img)[masked]
            dfeat = ops.gemm(da, rt.store.w16(iemb.img_linear.weight), b_mn=True, impl=rt.gemm_impl)
            ids = img_masks.long().contiguous()
            ops._call("b200u_embedding_scatter_add", P(dfeat), P(ids), R, R, C.c_longlong(0),
                      P(grad_buf(iemb.mask_embedding.weight)), n, D, C.c_longlong(0),
                      C.c_longlong(iemb.mask_embedding.weight.shape[0]))
        return None, None, None, None, None, None, None, None


class GatherFn(torch.autograd.Function):
    """torch.gather(torch.cat([txt_emb, img_emb], 1), 1, gather_index) (model/model.py:329-333)."""

    @staticmethod
    def forward(ctx, txt_emb, img_emb, gather_index):
        gather_index = _ids64(gather_index, "gather_index")
        ctx.gi = gather_index
        ctx.TR = (txt_emb.shape[1], img_emb.shape[1])
        return ops.gather_rows(txt_emb.contiguous(), img_emb.contiguous(), gather_index)

    @staticmethod
    def backward(ctx, dout):
        T, R = ctx.TR
        dtxt, dimg = ops.gather_rows_bwd(dout.contiguous(), ctx.gi, T, R)
        return dtxt, dimg, None


class PoolerFn(torch.autograd.Function):
    """BertPooler.forward (model/layer.py:179-185): tanh(dense(h[:, 0])) in fp32."""

    @staticmethod
    def forward(ctx, hidden, weight, bias):
        B, L, H = hidden.shape
        assert hidden.dtype == torch.bfloat16 and hidden.is_contiguous()
        pooled = torch.empty(B, H, device=hidden.device, dtype=torch.float32)
        ops._call("b200u_pooler_fwd", P(hidden), C.c_longlong(L * H), P(weight), P(bias), P(pooled), B, H)
        ctx.save_for_backward(hidden, pooled)
        ctx.params = (weight, bias)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        hidden, pooled = ctx.saved_tensors
        weight, bias = ctx.params
        B, L, H = hidden.shape
        dh = torch.zeros_like(hidden)
        dpooled = dpooled.contiguous().float()  # keep a reference until after the launch
        ops._call("b200u_pooler_bwd", P(dpooled), P(pooled), P(hidden),
                  C.c_longlong(L * H), P(weight), P(grad_buf(weight)), P(grad_buf(bias)), P(dh),
                  C.c_longlong(L * H), B, H)
        return dh, None, None


class SmallLinearFn(torch.autograd.Function):
    """x·Wᵀ + b for a handful of output classes, fp32 (model/meme_uniter.py:20, pretrain.py:62)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        B, K = x.shape
        Cn = weight.shape[0]
        x = x.contiguous().float()
        out = torch.empty(B, Cn, device=x.device, dtype=torch.float32)
        ops._call("b200u_linear_small_fwd", P(x), P(weight), P(bias), P(out), B, Cn, K)
        ctx.save_for_backward(x)
        ctx.params = (weight, bias)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        weight, bias = ctx.params
        B, K = x.shape
        Cn = weight.shape[0]
        dx = torch.empty_like(x)
        dout = dout.contiguous().float()  # keep a reference until after the launch
        gw = grad_buf(weight)
        gb = grad_buf(bias) if bias is not None else None
        ops._call("b200u_linear_small_bwd", P(dout), P(x), P(weight), P(dx), P(gw), P(gb), B, Cn, K)
        return dx, None, None


class LayerNormFn(torch.autograd.Function):
    """Standalone FusedLayerNorm (Apex replacement) on bf16 or fp32 CUDA tensors."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x = x.contiguous()
        y, mean, rstd = ops.layernorm_fwd(x, weight, bias, eps)
        ctx.save_for_backward(x, mean, rstd)
        ctx.params = (weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        weight, bias = ctx.params
        dx, _ = ops.layernorm_bwd(dy.contiguous().to(torch.bfloat16), x, mean, rstd, weight,
                                  grad_buf(weight), grad_buf(bias))
        return dx.to(x.dtype), None, None, None


def bce_with_logits(logits, labels, pos_weight=1.0, grad_scale=1.0, want_grad=True, dl_out=None):
    """Fused BCEWithLogitsLoss(pos_weight) mean loss, its gradient and sigmoid probabilities
    (train_template.py:64-65,98-99,117-118) in one launch. Returns (loss[1], dlogits[B], probs[B]).
    dl_out: optional contiguous f32 [B] destination for the gradient (a slice of a larger buffer)."""
    x = logits.reshape(-1).contiguous().float()
    y = labels.reshape(-1).contiguous().float()
    B = x.numel()
    loss = torch.empty(1, device=x.device, dtype=torch.float32)
    if dl_out is not None:
        assert dl_out.dtype == torch.float32 and dl_out.numel() == B and dl_out.is_contiguous()
        dl = dl_out
    else:
        dl = torch.empty(B, device=x.device, dtype=torch.float32) if want_grad else None
    probs = torch.empty(B, device=x.device, dtype=torch.float32)
    ops._call("b200u_bce_logits", P(x), P(y), float(pos_weight), float(grad_scale), P(loss), P(dl), P(probs), B)
    return loss, dl, probs


# ------------------------------------------------------------------------------------------------
# Generic nn.Linear on the tcgen05 GEMM (pretraining heads: model/layer.py:188-233,
# model/pretrain.py:19-47). Weight shadows come from the owning FlatStore when there is one (tied
# weights: word_embeddings / img_linear), else from a small per-parameter cache.
# ------------------------------------------------------------------------------------------------
_shadow_cache = {}


def shadow_for(p):
    from .flat import store_of
    st = store_of(p)
    if st is not None:
        st.refresh_shadow()
        return st.w16(p)
    key = id(p)
    hit = _shadow_cache.get(key)
    if hit is None or hit[0] != p._version or hit[1] != p.data_ptr():
        w16 = torch.empty(p.shape, device=p.device, dtype=torch.bfloat16)
        ops.cast_f32_to_bf16(p.detach().contiguous().view(-1), w16.view(-1))
        hit = (p._version, p.data_ptr(), w16)
        _shadow_cache[key] = hit
    return hit[2]


def _pad8(n):
    return (n + 7) // 8 * 8


class LinearFn(torch.autograd.Function):
    """y = x·Wᵀ + b (or x·W + b when `transposed`, model/pretrain.py:31) with optional fused GELU.
    x bf16 [n, K]; y bf16, or fp32 when out_f32 (logits)."""

    @staticmethod
    def forward(ctx, x, weight, bias, transposed, gelu, out_f32):
        assert x.dim() == 2
        x = x.to(torch.bfloat16).contiguous()
        n, K = x.shape
        w16 = shadow_for(weight)
        N = weight.shape[1] if transposed else weight.shape[0]
        ld = _pad8(N)
        if n == 0:
            ctx.empty = True
            ctx.shape = (n, K)
            return torch.zeros(0, N, device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
        ctx.empty = False
        u = None
        if gelu:
            ubuf = torch.empty(n, ld, device=x.device, dtype=torch.bfloat16)
            gbuf = torch.empty(n, ld, device=x.device, dtype=torch.bfloat16)
            ops.gemm(x, w16, b_mn=transposed, epilogue=_lib.EPI_BIAS_GELU, out=ubuf[:, :N], out2=gbuf[:, :N],
                     bias=bias)
            u, y = ubuf[:, :N], gbuf[:, :N]
        else:
            ybuf = torch.empty(n, ld, device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
            y = ybuf[:, :N]
            ops.gemm(x, w16, b_mn=transposed, epilogue=EPI_STORE_F32 if out_f32 else _lib.EPI_STORE, out=y,
                     bias=bias)
        ctx.save_for_backward(x, u)
        ctx.params = (weight, bias, w16)
        ctx.flags = (transposed, gelu)
        return y

    @staticmethod
    def backward(ctx, dy):
        if ctx.empty:
            return torch.zeros(ctx.shape, device=dy.device, dtype=torch.bfloat16), None, None, None, None, None
        x, u = ctx.saved_tensors
        weight, bias, w16 = ctx.params
        transposed, gelu = ctx.flags
        n, K = x.shape
        N = dy.shape[1]
        ld = _pad8(N)
        dyb = torch.zeros(n, ld, device=dy.device, dtype=torch.bfloat16)
        dyb[:, :N] = dy  # bf16, 16-byte aligned rows (N may be 28996)
        dyv = dyb[:, :N]
        if gelu:
            # d(pre-activation) = dy * gelu'(u); u lives in a buffer with the same padded stride
            ubuf = u._base if u._base is not None else u
            ops._call("b200u_dgelu_mul", P(dyb), P(ubuf), P(dyb), C.c_size_t(dyb.numel()))
        if bias is not None:
            ops.colsum_accum(dyb, grad_buf(bias)) if ld == N else grad_buf(bias).add_(dyv.float().sum(0))
        gw = grad_buf(weight)
        if transposed:
            # W [K, N]: dW += xᵀ·dy ; dx = dy·Wᵀ
            ops.gemm(x, dyv, a_mn=True, b_mn=True, epilogue=EPI_ATOMIC_F32, out=gw)
            dx = ops.gemm(dyv, w16)
        else:
            # W [N, K]: dW += dyᵀ·x ; dx = dy·W
            ops.gemm(dyv, x, a_mn=True, b_mn=True, epilogue=EPI_ATOMIC_F32, out=gw)
            dx = ops.gemm(dyv, w16, b_mn=True)
        return dx, None, None, None, None, None


def linear(x, weight, bias=None, transposed=False, gelu=False, out_f32=False):
    return LinearFn.apply(x, weight, bias, transposed, gelu, out_f32)


class VocabCrossEntropyFn(torch.autograd.Function):
    """F.cross_entropy(x·Wᵀ + b, targets, reduction='none') WITHOUT the [n, vocab] logits (the MLM decoder tied to
    the word embeddings followed by the loss, model/layer.py:204-221 + model/pretrain.py:97-98).

    forward: one GEMM whose epilogue reduces every 128-column tile of a row to (max, sum exp) and picks out the
    target logit (EPI_CE_STATS), then b200u_ce_finish -> lse, loss. backward: the GEMM is run again with the
    epilogue that turns the recomputed logits into d logits = (softmax - onehot) * d loss as the bf16 operand
    (EPI_CE_GRAD) of the decoder's dgrad / wgrad GEMMs and the bias column sums. fp32 logits are never stored."""

    @staticmethod
    def forward(ctx, x, weight, bias, targets):
        assert x.dim() == 2
        x = x.to(torch.bfloat16).contiguous()
        n, K = x.shape
        N = weight.shape[0]
        if n == 0:
            ctx.empty = True
            ctx.shape = (n, K)
            return torch.zeros(0, device=x.device, dtype=torch.float32)
        ctx.empty = False
        w16 = shadow_for(weight)
        targets = _ids64(targets, "targets")
        if bias is None:
            bias = torch.zeros(N, device=x.device, dtype=torch.float32)
            ctx.has_bias = False
        else:
            ctx.has_bias = True
        nt = (N + 127) // 128
        partial = torch.empty(n, nt, 2, device=x.device, dtype=torch.float32)
        tlogit = torch.zeros(n, device=x.device, dtype=torch.float32)
        ops.gemm(x, w16, epilogue=_lib.EPI_CE_STATS, bias=bias, ce=dict(target=targets, partial=partial, tlogit=tlogit))
        lse, loss = ops.ce_finish(partial, tlogit)
        ctx.save_for_backward(x, targets, lse)
        ctx.params = (weight, bias, w16)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        if ctx.empty:
            return torch.zeros(ctx.shape, device=dloss.device, dtype=torch.bfloat16), None, None, None
        x, targets, lse = ctx.saved_tensors
        weight, bias, w16 = ctx.params
        n, K = x.shape
        N = weight.shape[0]
        ld = _pad8(N)
        dyb = torch.zeros(n, ld, device=x.device, dtype=torch.bfloat16)   # 16-byte aligned rows, zero pad columns
        dyv = dyb[:, :N]
        ops.gemm(x, w16, epilogue=_lib.EPI_CE_GRAD, out=dyv, bias=bias,
                 ce=dict(target=targets, lse=lse, scale=dloss.float().contiguous()))
        if ctx.has_bias:
            ops.colsum_accum(dyb, grad_buf(bias)) if ld == N else grad_buf(bias).add_(dyv.float().sum(0))
        ops.gemm(dyv, x, a_mn=True, b_mn=True, epilogue=EPI_ATOMIC_F32, out=grad_buf(weight))
        dx = ops.gemm(dyv, w16, b_mn=True)
        return dx, None, None, None


def vocab_cross_entropy(x, weight, bias, targets):
    return VocabCrossEntropyFn.apply(x, weight, bias, targets)
