"""Synthetic batches of the BASELINE shapes: what bench.py, smoke() and the tools feed
the hot path with. Keys follow the reference collates (data/meme_dataset.py:204-212 for fine-tuning,
data/pretrain_{mlm,mrfr,itm}.py for pretraining); index / mask construction goes through the
package's own bit-exact `utils.get_gather_index` / `get_attention_mask`.

tests/test_host_logic.py checks that these generators return exactly what the oracle's generators
(the ones the golden fixtures were made with) return for the same seeds.
"""
import torch

from ..utils.utils import get_attention_mask, get_gather_index


def synth_batch(B, T, R, seed=1234, variable=False, img_dim=2048, vocab=28996, min_txt=8, min_bb=36):
    """One fine-tuning micro-batch on the host (CPU tensors).

    input_ids: [CLS]=101 ... [SEP]=102, pad 0 to width T (train_uniter.py:125 padding='max_length');
    position_ids: arange(T) per sample (meme_dataset.py:179); img_feat = relu(N(0,1)) fp32 with rows
    >= num_bb zeroed (pad_sequence, meme_dataset.py:161); img_pos_feat = [x1,y1,x2,y2,w,h,w*h]
    (dataset_template.py:106-113); attn_mask / gather_index from utils/utils.py:111-125;
    labels ~ Bernoulli(0.36)."""
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


def synth_pretrain_batch(B, T, R, seed=77, img_dim=2048, vocab=28996, label_dim=1601, min_txt=8, min_bb=36,
                         variable=True):
    """Multi-task pretraining batch: the union of the keys the MLM / MRFR / MRC / ITM collates produce
    (data/pretrain_mlm.py:110-127, pretrain_mrfr.py:29-35) plus an `ot_inputs` dict with the semantics
    model/pretrain.py:169-190 expects. Mask probability 0.15, at least one masked
    token / region per sample."""
    b = synth_batch(B, T, R, seed=seed, variable=variable, img_dim=img_dim, vocab=vocab, min_txt=min_txt,
                    min_bb=min_bb)
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
