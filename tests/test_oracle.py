"""CPU: the oracle restatement (oracle/uniter_oracle.py) against golden vectors produced by the
UNMODIFIED reference modules (oracle/make_golden.py). This is the parity pin of the oracle."""
import os

import numpy as np
import torch

from oracle import uniter_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_index_and_mask_bit_exact(golden_dir):
    g = _load(golden_dir, "index_mask.npz")
    i = 0
    while "c%d_T" % i in g:
        tl, nb, T = g["c%d_txt_lens" % i].tolist(), g["c%d_num_bbs" % i].tolist(), int(g["c%d_T" % i])
        am = O.get_attention_mask(tl, nb)
        gi = O.get_gather_index(tl, nb, len(tl), T, am.shape[1])
        assert am.dtype == torch.float32 and gi.dtype == torch.int64
        assert np.array_equal(am.numpy(), g["c%d_attn_mask" % i])
        assert np.array_equal(gi.numpy(), g["c%d_gather_index" % i])
        i += 1
    assert i == 4


def _tiny(golden_dir):
    g = _load(golden_dir, "tiny_meme_uniter.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    b = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in.")}
    return g, sd, b


def test_tiny_forward_matches_reference(golden_dir):
    from oracle.make_golden import TINY
    g, sd, b = _tiny(golden_dir)
    kw = dict(input_ids=b["input_ids"], position_ids=b["position_ids"], img_feat=b["img_feat"],
              img_pos_feat=b["img_pos_feat"], attention_mask=b["attn_mask"], gather_index=b["gather_index"])
    with torch.no_grad():
        emb, _, _ = O.uniter_embeddings(sd, "uniter_model.", b["input_ids"], b["position_ids"], b["img_feat"],
                                        b["img_pos_feat"], b["gather_index"])
        layers = O.uniter_forward(sd, TINY, pre="uniter_model.", **kw)
        logits = O.meme_uniter_forward(sd, TINY, **kw)
        loss = O.bce_loss(logits, b["labels"], 1.8)
    assert np.array_equal(emb.numpy(), g["embedding_output"])  # same ops, same order -> bitwise
    np.testing.assert_allclose(layers[0].numpy(), g["layer0_out"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(layers[1].numpy(), g["layer1_out"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(loss.numpy(), g["loss"], rtol=1e-6)


def test_tiny_gradients_match_reference(golden_dir):
    from oracle.make_golden import TINY
    g, sd, b = _tiny(golden_dir)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    logits = O.meme_uniter_forward(sd, TINY, input_ids=b["input_ids"], position_ids=b["position_ids"],
                                   img_feat=b["img_feat"], img_pos_feat=b["img_pos_feat"],
                                   attention_mask=b["attn_mask"], gather_index=b["gather_index"])
    O.bce_loss(logits, b["labels"], 1.8).backward()
    checked = 0
    for k in g.files:
        if not k.startswith("grad."):
            continue
        got = sd[k[5:]].grad
        assert got is not None, k
        want = g[k]
        np.testing.assert_allclose(got.numpy(), want, rtol=1e-4, atol=1e-7, err_msg=k)
        checked += 1
    assert checked >= 40
    assert sd["uniter_model.img_embeddings.mask_embedding.weight"].grad is None


def test_ot_matches_reference(golden_dir):
    g = _load(golden_dir, "ot.npz")
    txt, img = torch.from_numpy(g["txt"]), torch.from_numpy(g["img"])
    txt_pad, img_pad = torch.from_numpy(g["txt_pad"]), torch.from_numpy(g["img_pad"])
    dist, T, _ = O.optimal_transport_dist(txt, img, txt_pad, img_pad)
    np.testing.assert_allclose(O.cost_matrix_cosine(txt, img).numpy(), g["cost"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(T.numpy(), g["T"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(dist.numpy(), g["dist"], rtol=1e-5)
    # marginals of the transport plan converge to 1/len on the valid block
    np.testing.assert_allclose(T[0].sum().item(), 1.0, rtol=1e-3)


def test_adam_l2_step_matches_torch():
    """oracle step semantics vs torch.optim.Adam + clip_grad_norm_ + grad/accum
    (train_template.py:89-107, utils/optim_utils.py:16-46)."""
    torch.manual_seed(0)
    names = ["a.weight", "a.bias", "x.LayerNorm.weight", "img_layer_norm.weight"]
    ps = [torch.randn(7, 5), torch.randn(7), torch.randn(5), torch.randn(5)]
    gs = [torch.randn_like(p) * 3 for p in ps]
    tp = [torch.nn.Parameter(p.clone()) for p in ps]
    groups = [{"params": [p for n, p in zip(names, tp) if not O.is_no_decay(n)], "weight_decay": 1e-3},
              {"params": [p for n, p in zip(names, tp) if O.is_no_decay(n)], "weight_decay": 0.0}]
    assert [O.is_no_decay(n) for n in names] == [False, True, True, False]  # img_layer_norm IS decayed
    opt = torch.optim.Adam(groups, lr=3e-3, betas=(0.9, 0.999))
    state = {}
    cur = ps
    for _ in range(3):
        for p, g in zip(tp, gs):
            p.grad = g.clone() / 2
        torch.nn.utils.clip_grad_norm_(tp, 5)
        opt.step()
        cur, _ = O.adam_l2_step(cur, gs, state, names, 3e-3, 1e-3, 2, 5)
    for a, b in zip(cur, tp):
        np.testing.assert_allclose(a.numpy(), b.detach().numpy(), rtol=1e-5, atol=1e-7)


def _pretrain_golden(golden_dir):
    g = _load(golden_dir, "tiny_pretrain.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    b = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in.")}
    oi = {k[3:]: (torch.from_numpy(g[k]) if g[k].ndim else int(g[k])) for k in g.files if k.startswith("ot.")}
    b["ot_inputs"] = oi
    return g, sd, b


def test_pretraining_heads_match_reference(golden_dir):
    """oracle pretrain_forward vs the unmodified UniterForPretraining (model/pretrain.py:65-233)."""
    from oracle.make_golden import TINY
    g, sd, b = _pretrain_golden(golden_dir)
    with torch.no_grad():
        np.testing.assert_allclose(O.pretrain_forward(sd, TINY, b, "mlm").numpy(), g["loss.mlm"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(O.pretrain_forward(sd, TINY, b, "mlm", False).numpy(), g["scores.mlm"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(O.pretrain_forward(sd, TINY, b, "mrfr").numpy(), g["loss.mrfr"], rtol=1e-4, atol=1e-6)
        itm, ot = O.pretrain_forward(sd, TINY, b, "itm")
        np.testing.assert_allclose(itm.numpy(), g["loss.itm"], rtol=1e-5, atol=1e-6)
        assert ot.shape == (4,) and torch.isfinite(ot).all()
        np.testing.assert_allclose(O.pretrain_forward(sd, TINY, b, "mrc-kl").numpy(), g["loss.mrc-kl"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(O.pretrain_forward(sd, TINY, b, "mrc").numpy(), g["loss.mrc"], rtol=1e-5, atol=1e-6)
