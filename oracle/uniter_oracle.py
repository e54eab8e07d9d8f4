"""TEST INFRASTRUCTURE ONLY — plain-PyTorch (CPU, fp32) restatement of the reference hot path.

Every function cites the reference lines it restates (paths relative to the reference repo
Nithin-Holla/meme_challenge). The restatement is *functional*: it takes a state_dict with the
reference's key names and never imports the reference, so it runs on the GPU box where
/root/reference does not exist. Parity pin: oracle/make_golden.py runs the UNMODIFIED reference
modules (apex FusedLayerNorm aliased to torch.nn.LayerNorm, ot.trace patched for torch>=2) in the
build container and stores their outputs under tests/golden/; tests/test_oracle.py checks this
file against those vectors. Third-party arithmetic outside the reference tree: apex FusedLayerNorm
(un-pinned master, README.md:10-15) == (x-mean)/sqrt(biased_var+eps)*w+b, restated with
F.layer_norm; everything else is stock torch.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# utils/utils.py:111-125 — index / mask construction (exact loops of the reference)
# ----------------------------------------------------------------------------------------------
def get_gather_index(txt_lens, num_bbs, batch_size, max_len, out_size):
    """utils/utils.py:111-117."""
    assert len(txt_lens) == len(num_bbs) == batch_size
    gather_index = torch.arange(0, out_size, dtype=torch.long).unsqueeze(0).repeat(batch_size, 1)
    for i, (tl, nbb) in enumerate(zip(txt_lens, num_bbs)):
        gather_index.data[i, tl:tl + nbb] = torch.arange(max_len, max_len + nbb, dtype=torch.long).data
    return gather_index


def get_attention_mask(text_len, img_len):
    """utils/utils.py:120-125."""
    attn_mask = []
    for i in range(len(text_len)):
        attn_mask.append(torch.ones(text_len[i] + img_len[i]))
    return torch.nn.utils.rnn.pad_sequence(attn_mask, batch_first=True, padding_value=0)


# ----------------------------------------------------------------------------------------------
# model/layer.py — BERT layers
# ----------------------------------------------------------------------------------------------
def gelu(x):
    """model/layer.py:31-37."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b, eps=1e-12):
    """apex FusedLayerNorm(H, eps=1e-12) call sites model/model.py:229,252,253,258, layer.py:108,149."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def self_attention(sd, pre, h, ext_mask, heads, p_attn=0.0, training=False):
    """model/layer.py:75-101."""
    B, L, H = h.shape
    d = H // heads

    def split(x):
        return x.view(B, L, heads, d).permute(0, 2, 1, 3)
    q = split(F.linear(h, sd[pre + "query.weight"], sd[pre + "query.bias"]))
    k = split(F.linear(h, sd[pre + "key.weight"], sd[pre + "key.bias"]))
    v = split(F.linear(h, sd[pre + "value.weight"], sd[pre + "value.bias"]))
    scores = torch.matmul(q, k.transpose(-1, -2))
    scores = scores / math.sqrt(d)
    scores = scores + ext_mask
    probs = torch.softmax(scores, dim=-1)
    probs = F.dropout(probs, p_attn, training)
    ctx = torch.matmul(probs, v)
    return ctx.permute(0, 2, 1, 3).contiguous().view(B, L, H)


def bert_layer(sd, pre, h, ext_mask, heads, p_hidden=0.0, p_attn=0.0, training=False):
    """model/layer.py:166-170 (BertAttention 125-127, BertSelfOutput 111-115, BertIntermediate
    139-142, BertOutput 152-156)."""
    a = self_attention(sd, pre + "attention.self.", h, ext_mask, heads, p_attn, training)
    a = F.linear(a, sd[pre + "attention.output.dense.weight"], sd[pre + "attention.output.dense.bias"])
    a = F.dropout(a, p_hidden, training)
    a = layer_norm(a + h, sd[pre + "attention.output.LayerNorm.weight"], sd[pre + "attention.output.LayerNorm.bias"])
    i = gelu(F.linear(a, sd[pre + "intermediate.dense.weight"], sd[pre + "intermediate.dense.bias"]))
    o = F.linear(i, sd[pre + "output.dense.weight"], sd[pre + "output.dense.bias"])
    o = F.dropout(o, p_hidden, training)
    return layer_norm(o + a, sd[pre + "output.LayerNorm.weight"], sd[pre + "output.LayerNorm.bias"])


# ----------------------------------------------------------------------------------------------
# model/model.py — embeddings, gather, encoder
# ----------------------------------------------------------------------------------------------
def text_embeddings(sd, pre, input_ids, position_ids, token_type_ids=None):
    """model/model.py:232-245 (dropout off)."""
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)
    e = (F.embedding(input_ids, sd[pre + "word_embeddings.weight"])
         + F.embedding(position_ids, sd[pre + "position_embeddings.weight"])
         + F.embedding(token_type_ids, sd[pre + "token_type_embeddings.weight"]))
    return layer_norm(e, sd[pre + "LayerNorm.weight"], sd[pre + "LayerNorm.bias"])


def image_embeddings(sd, pre, img_feat, img_pos_feat, type_embeddings, img_masks=None):
    """model/model.py:261-272 (dropout off)."""
    if img_masks is not None:
        w = sd[pre + "mask_embedding.weight"].clone()
        w[0, :] = 0
        img_feat = img_feat + F.embedding(img_masks.long(), w)
    ti = layer_norm(F.linear(img_feat, sd[pre + "img_linear.weight"], sd[pre + "img_linear.bias"]),
                    sd[pre + "img_layer_norm.weight"], sd[pre + "img_layer_norm.bias"])
    tp = layer_norm(F.linear(img_pos_feat, sd[pre + "pos_linear.weight"], sd[pre + "pos_linear.bias"]),
                    sd[pre + "pos_layer_norm.weight"], sd[pre + "pos_layer_norm.bias"])
    return layer_norm(ti + tp + type_embeddings, sd[pre + "LayerNorm.weight"], sd[pre + "LayerNorm.bias"])


def uniter_embeddings(sd, pre, input_ids, position_ids, img_feat, img_pos_feat, gather_index,
                      img_masks=None, txt_type_ids=None, img_type_ids=None):
    """model/model.py:305-334 (_compute_txt/img/img_txt_embeddings)."""
    txt = None
    if input_ids is not None:
        txt = text_embeddings(sd, pre + "embeddings.", input_ids, position_ids, txt_type_ids)
    img = None
    if img_feat is not None:
        if img_type_ids is None:
            img_type_ids = torch.ones_like(img_feat[:, :, 0].long())
        te = F.embedding(img_type_ids, sd[pre + "embeddings.token_type_embeddings.weight"])
        img = image_embeddings(sd, pre + "img_embeddings.", img_feat, img_pos_feat, te, img_masks)
    if txt is None:
        return img, None, img
    if img is None:
        return txt, txt, None
    H = txt.shape[-1]
    gi = gather_index.unsqueeze(-1).expand(-1, -1, H)
    return torch.gather(torch.cat([txt, img], dim=1), dim=1, index=gi), txt, img


def uniter_forward(sd, cfg, input_ids, position_ids, img_feat, img_pos_feat, attention_mask,
                   gather_index=None, img_masks=None, output_all_encoded_layers=True,
                   txt_type_ids=None, img_type_ids=None, pre="", p_hidden=0.0, p_attn=0.0, training=False):
    """model/model.py:336-367 (UniterModel.forward) + UniterEncoder.forward 282-292."""
    ext = attention_mask.unsqueeze(1).unsqueeze(2).to(torch.float32)
    ext = (1.0 - ext) * -10000.0
    h, _, _ = uniter_embeddings(sd, pre, input_ids, position_ids, img_feat, img_pos_feat, gather_index,
                                img_masks, txt_type_ids, img_type_ids)
    outs = []
    for l in range(cfg["num_hidden_layers"]):
        h = bert_layer(sd, "%sencoder.layer.%d." % (pre, l), h, ext, cfg["num_attention_heads"],
                       p_hidden, p_attn, training)
        if output_all_encoded_layers:
            outs.append(h)
    return outs if output_all_encoded_layers else h


def pooler(sd, pre, h):
    """model/layer.py:179-185."""
    return torch.tanh(F.linear(h[:, 0], sd[pre + "dense.weight"], sd[pre + "dense.bias"]))


def meme_uniter_forward(sd, cfg, **kw):
    """model/meme_uniter.py:17-21 (state_dict keys uniter_model.* + linear.*)."""
    kw.setdefault("output_all_encoded_layers", False)
    h = uniter_forward(sd, cfg, pre="uniter_model.", **kw)
    p = pooler(sd, "uniter_model.pooler.", h)
    return F.linear(p, sd["linear.weight"], sd["linear.bias"])


def bce_loss(logits, labels, pos_wt):
    """train_template.py:64-65,98-99: BCEWithLogitsLoss(pos_weight=[pos_wt]) on preds.squeeze(1)."""
    return F.binary_cross_entropy_with_logits(logits.squeeze(1), labels.float(),
                                              pos_weight=torch.tensor([pos_wt], device=logits.device))


# ----------------------------------------------------------------------------------------------
# train_template.py:89-109 + utils/optim_utils.py:16-46 — step semantics
# ----------------------------------------------------------------------------------------------
NO_DECAY = ['bias', 'LayerNorm.bias', 'LayerNorm.weight']


def is_no_decay(name):
    """utils/optim_utils.py:16-24 (case-sensitive substring match)."""
    return any(nd in name for nd in NO_DECAY)


def adam_l2_step(params, grads, state, names, lr, weight_decay, accum, max_grad_norm,
                 beta1=0.9, beta2=0.999, eps=1e-8):
    """One optimizer step as train_template.py:101-107 does it: grads /= accum (89-92), global
    clip_grad_norm_(max_grad_norm) (104), torch.optim.Adam with L2 decay on the decay group."""
    grads = [g / accum for g in grads]
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_grad_norm / (total + 1e-6), max=1.0)
    grads = [g * coef for g in grads]
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    out = []
    for i, (p, g, n) in enumerate(zip(params, grads, names)):
        wd = 0.0 if is_no_decay(n) else weight_decay
        g = g + wd * p
        m = state.setdefault(("m", i), torch.zeros_like(p))
        v = state.setdefault(("v", i), torch.zeros_like(p))
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / math.sqrt(1 - beta2 ** t)).add_(eps)
        out.append(p - (lr / (1 - beta1 ** t)) * m / denom)
    return out, float(total)


# ----------------------------------------------------------------------------------------------
# model/ot.py — IPOT optimal transport
# ----------------------------------------------------------------------------------------------
def cost_matrix_cosine(x, y, eps=1e-5):
    """model/ot.py:11-21."""
    assert x.dim() == y.dim()
    assert x.size(0) == y.size(0)
    assert x.size(2) == y.size(2)
    x_norm = F.normalize(x, p=2, dim=-1, eps=eps)
    y_norm = F.normalize(y, p=2, dim=-1, eps=eps)
    return 1 - x_norm.matmul(y_norm.transpose(1, 2))


def trace(x):
    """model/ot.py:24-32 (same elements, summed along the diagonal; the reference's uint8 mask
    raises on torch>=2, see SURVEY.md §8c shim 2)."""
    b, m, n = x.size()
    assert m == n
    return torch.diagonal(x, dim1=-2, dim2=-1).sum(-1)


@torch.no_grad()
def ipot(C, x_len, x_pad, y_len, y_pad, joint_pad, beta, iteration, k):
    """model/ot.py:35-66. [B,M,N],[B],[B,M],[B],[B,N],[B,M,N] -> T [B,N,M]."""
    b, m, n = C.size()
    sigma = torch.ones(b, m, dtype=C.dtype, device=C.device) / x_len.unsqueeze(1)
    T = torch.ones(b, n, m, dtype=C.dtype, device=C.device)
    A = torch.exp(-C.transpose(1, 2) / beta)
    sigma.masked_fill_(x_pad, 0)
    joint_pad = joint_pad.transpose(1, 2)
    T.masked_fill_(joint_pad, 0)
    A.masked_fill_(joint_pad, 0)
    x_len = x_len.unsqueeze(1).unsqueeze(2)
    y_len = y_len.unsqueeze(1).unsqueeze(2)
    x_mask = (x_pad.to(C.dtype) * 1e4).unsqueeze(1)
    y_mask = (y_pad.to(C.dtype) * 1e4).unsqueeze(1)
    for _ in range(iteration):
        Q = A * T
        sigma = sigma.view(b, m, 1)
        for _ in range(k):
            delta = 1 / (y_len * Q.matmul(sigma).view(b, 1, n) + y_mask)
            sigma = 1 / (x_len * delta.matmul(Q) + x_mask)
        T = delta.view(b, n, 1) * Q * sigma
    T.masked_fill_(joint_pad, 0)
    return T


def optimal_transport_dist(txt_emb, img_emb, txt_pad, img_pad, beta=0.5, iteration=50, k=1):
    """model/ot.py:69-85."""
    cost = cost_matrix_cosine(txt_emb, img_emb)
    joint_pad = txt_pad.unsqueeze(-1) | img_pad.unsqueeze(-2)
    cost = cost.masked_fill(joint_pad, 0)
    txt_len = (txt_pad.size(1) - txt_pad.sum(dim=1, keepdim=False)).to(dtype=cost.dtype)
    img_len = (img_pad.size(1) - img_pad.sum(dim=1, keepdim=False)).to(dtype=cost.dtype)
    T = ipot(cost.detach(), txt_len, txt_pad, img_len, img_pad, joint_pad, beta, iteration, k)
    return trace(cost.matmul(T.detach())), T, cost


# ----------------------------------------------------------------------------------------------
# Synthetic batches: shared by tests, smoke() and bench.py
# ----------------------------------------------------------------------------------------------
def synth_batch(B, T, R, seed=1234, variable=False, img_dim=2048, vocab=28996, min_txt=8, min_bb=36):
    g = torch.Generator().manual_seed(seed)
    if variable:
        txt_lens = torch.randint(min(min_txt, T), T + 1, (B,), generator=g).tolist()
        num_bbs = torch.randint(min(min_bb, R), R + 1, (B,), generator=g).tolist()
    else:
        txt_lens, num_bbs = [T] * B, [R] * B
    maxR = max(num_bbs)
    input_ids = torch.zeros(B, T, dtype=torch.long)
    for i, tl in enumerate(txt_lens):
        lo = min(1000, vocab - 1)
        input_ids[i, :tl] = torch.randint(lo, vocab, (tl,), generator=g)
        input_ids[i, 0] = min(101, vocab - 1)
        input_ids[i, tl - 1] = min(102, vocab - 1)
    position_ids = torch.arange(T, dtype=torch.long).unsqueeze(0).repeat(B, 1)
    img_feat = torch.relu(torch.randn(B, maxR, img_dim, generator=g))
    xy = torch.rand(B, maxR, 2, generator=g) * 0.7
    wh = torch.rand(B, maxR, 2, generator=g) * 0.25 + 0.05
    pos = torch.cat([xy, xy + wh, wh, wh[..., :1] * wh[..., 1:]], dim=-1)
    for i, nb in enumerate(num_bbs):
        img_feat[i, nb:] = 0
        pos[i, nb:] = 0
    attn = get_attention_mask(txt_lens, num_bbs)
    L = attn.shape[1]
    gi = get_gather_index(txt_lens, num_bbs, B, T, L)
    labels = (torch.rand(B, generator=g) < 0.36).long()
    return dict(input_ids=input_ids, position_ids=position_ids, img_feat=img_feat, img_pos_feat=pos,
                attn_mask=attn, gather_index=gi, labels=labels, txt_lens=txt_lens, num_bbs=num_bbs)


# ----------------------------------------------------------------------------------------------
# model/pretrain.py + model/layer.py:188-233 — pretraining heads (state_dict keys of
# UniterForPretraining: uniter.*, cls.predictions.*, feat_regress.*, region_classifier.*, itm_output.*)
# ----------------------------------------------------------------------------------------------
def _masked_hidden(hidden, mask):
    """model/pretrain.py:129-133."""
    mask = mask.unsqueeze(-1).expand_as(hidden)
    return hidden[mask].contiguous().view(-1, hidden.size(-1))


def _transform(sd, pre_dense, pre_ln, x):
    """dense -> gelu -> LayerNorm (layer.py:196-200; pretrain.py:23-25, 40-42)."""
    h = gelu(F.linear(x, sd[pre_dense + "weight"], sd[pre_dense + "bias"]))
    return layer_norm(h, sd[pre_ln + "weight"], sd[pre_ln + "bias"])


def pretrain_forward(sd, cfg, batch, task, compute_loss=True):
    """UniterForPretraining.forward (model/pretrain.py:65-233), eval mode."""
    kw = dict(input_ids=batch["input_ids"], position_ids=batch["position_ids"], img_feat=batch["img_feat"],
              img_pos_feat=batch["img_pos_feat"], attention_mask=batch["attn_masks"],
              gather_index=batch["gather_index"], output_all_encoded_layers=False, pre="uniter.")
    if task == "mlm":
        seq = uniter_forward(sd, cfg, **kw)[:, :batch["input_ids"].size(1), :]
        lab = batch["txt_labels"]
        h = _transform(sd, "cls.predictions.transform.dense.", "cls.predictions.transform.LayerNorm.",
                       _masked_hidden(seq, lab != -1))
        scores = F.linear(h, sd["cls.predictions.decoder.weight"]) + sd["cls.predictions.bias"]
        return F.cross_entropy(scores, lab[lab != -1], reduction="none") if compute_loss else scores
    if task == "mrfr":
        seq = uniter_forward(sd, cfg, img_masks=batch["img_masks"], **kw)
        h = _transform(sd, "feat_regress.net.0.", "feat_regress.net.2.", _masked_hidden(seq, batch["img_mask_tgt"]))
        pred = F.linear(h, sd["feat_regress.weight"].t(), sd["feat_regress.bias"])
        return F.mse_loss(pred, batch["feat_targets"], reduction="none") if compute_loss else pred
    if task == "itm":
        seq = uniter_forward(sd, cfg, **kw)
        scores = F.linear(pooler(sd, "uniter.pooler.", seq), sd["itm_output.weight"], sd["itm_output.bias"])
        ot = None
        if batch.get("ot_inputs") is not None:
            oi = batch["ot_inputs"]
            b, tl, il = seq.size(0), batch["input_ids"].size(1), batch["img_feat"].size(1)
            max_l = max(oi["scatter_max"] + 1, tl + il)
            idx = oi["ot_scatter"].unsqueeze(-1).expand_as(seq)
            ctx = torch.zeros(b, max_l, seq.size(-1), dtype=seq.dtype).scatter_(dim=1, index=idx, src=seq)
            ot, _, _ = optimal_transport_dist(ctx[:, :tl].float(), ctx[:, tl:tl + il].float(), oi["txt_pad"], oi["img_pad"])
        loss = F.cross_entropy(scores, batch["targets"], reduction="none") if compute_loss else scores
        return loss, ot
    if task.startswith("mrc"):
        seq = uniter_forward(sd, cfg, img_masks=batch["img_masks"], **kw)
        h = _transform(sd, "region_classifier.net.0.", "region_classifier.net.2.",
                       _masked_hidden(seq, batch["img_mask_tgt"]))
        pred = F.linear(h, sd["region_classifier.net.3.weight"], sd["region_classifier.net.3.bias"])
        if not compute_loss:
            return pred
        if "kl" in task:
            return F.kl_div(F.log_softmax(pred, dim=-1), batch["label_targets"], reduction="none")
        tgt = torch.max(batch["label_targets"][:, 1:], dim=-1)[1] + 1
        return F.cross_entropy(pred, tgt, ignore_index=0, reduction="none")
    raise ValueError("invalid task")


def synth_pretrain_batch(B, T, R, seed=77, img_dim=2048, vocab=28996, label_dim=1601, min_txt=8, min_bb=36):
    """Synthetic multi-task batch with the keys of data/pretrain_{mlm,mrfr,itm}.py collates and an
    ot_inputs dict whose semantics follow model/pretrain.py:169-190."""
    b = synth_batch(B, T, R, seed=seed, variable=True, img_dim=img_dim, vocab=vocab, min_txt=min_txt, min_bb=min_bb)
    g = torch.Generator().manual_seed(seed + 1)
    tl, nb = b["txt_lens"], b["num_bbs"]
    L = b["attn_mask"].shape[1]
    maxR = b["img_feat"].shape[1]
    txt_labels = torch.full((B, T), -1, dtype=torch.long)
    img_masks = torch.zeros(B, maxR, dtype=torch.bool)
    img_mask_tgt = torch.zeros(B, L, dtype=torch.bool)
    ot_scatter = torch.full((B, L), T + maxR, dtype=torch.long)
    for i in range(B):
        m = torch.rand(tl[i], generator=g) < 0.15
        m[int(torch.randint(0, tl[i], (1,), generator=g))] = True
        txt_labels[i, :tl[i]][m] = b["input_ids"][i, :tl[i]][m]
        r = torch.rand(nb[i], generator=g) < 0.15
        r[int(torch.randint(0, nb[i], (1,), generator=g))] = True
        img_masks[i, :nb[i]] = r
        img_mask_tgt[i, tl[i]:tl[i] + nb[i]] = r
        ot_scatter[i, :tl[i]] = torch.arange(tl[i])
        ot_scatter[i, tl[i]:tl[i] + nb[i]] = T + torch.arange(nb[i])
    feat_targets = b["img_feat"][img_masks].clone()
    label_targets = torch.softmax(torch.randn(int(img_masks.sum()), label_dim, generator=g), -1)
    out = dict(b)
    out.update(attn_masks=b["attn_mask"], txt_labels=txt_labels, img_masks=img_masks, img_mask_tgt=img_mask_tgt,
               feat_targets=feat_targets, label_targets=label_targets,
               targets=(torch.rand(B, generator=g) < 0.5).long(),
               ot_inputs=dict(ot_scatter=ot_scatter, scatter_max=T + maxR,
                              txt_pad=torch.arange(T).unsqueeze(0) >= torch.tensor(tl).unsqueeze(1),
                              img_pad=torch.arange(maxR).unsqueeze(0) >= torch.tensor(nb).unsqueeze(1)))
    return out
