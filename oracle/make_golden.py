"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules from /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
Shims: apex.normalization.fused_layer_norm.FusedLayerNorm -> torch.nn.LayerNorm
(apex is not installed; same arithmetic), and ot.trace -> diagonal sum (the reference's uint8
masked_select raises on torch>=2). Nothing else of the reference is touched.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("B200U_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    apex = types.ModuleType("apex")
    norm = types.ModuleType("apex.normalization")
    fln = types.ModuleType("apex.normalization.fused_layer_norm")
    fln.FusedLayerNorm = torch.nn.LayerNorm
    sys.modules["apex"] = apex
    sys.modules["apex.normalization"] = norm
    sys.modules["apex.normalization.fused_layer_norm"] = fln
    sys.path.insert(0, REF)
    import model.model as rmodel
    import model.meme_uniter as rmeme
    import model.ot as rot
    import utils.utils as rutils
    rot.trace = lambda x: torch.diagonal(x, dim1=-2, dim2=-1).sum(-1)
    return rmodel, rmeme, rot, rutils


TINY = dict(vocab_size=120, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
            intermediate_size=256, hidden_act="gelu", hidden_dropout_prob=0.1,
            attention_probs_dropout_prob=0.1, max_position_embeddings=64, type_vocab_size=2,
            initializer_range=0.02)
IMG_DIM = 64
LABEL_DIM = 24


def main():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle.uniter_oracle import synth_batch
    rmodel, rmeme, rot, rutils = import_reference()
    os.makedirs(OUT, exist_ok=True)

    # ---- 1. index / mask construction on ragged inputs (utils/utils.py:111-125)
    cases = [([5, 3, 8], [4, 7, 2], 8), ([1], [1], 1), ([40] * 4, [36] * 4, 40), ([8, 64, 33], [100, 36, 77], 64)]
    g = {}
    for i, (tl, nb, T) in enumerate(cases):
        am = rutils.get_attention_mask(tl, nb)
        gi = rutils.get_gather_index(tl, nb, len(tl), T, am.shape[1])
        g["c%d_txt_lens" % i] = np.array(tl); g["c%d_num_bbs" % i] = np.array(nb); g["c%d_T" % i] = np.array(T)
        g["c%d_attn_mask" % i] = am.numpy(); g["c%d_gather_index" % i] = gi.numpy()
    np.savez_compressed(os.path.join(OUT, "index_mask.npz"), **g)

    # ---- 2. tiny MemeUniter: weights, ragged batch, logits / loss / grads, eval mode
    torch.manual_seed(0)
    cfg = rmodel.UniterConfig.from_dict(TINY)
    um = rmodel.UniterModel(cfg, IMG_DIM)
    m = rmeme.MemeUniter(um, cfg.hidden_size, 1)
    # make LayerNorm affine params and biases non-trivial so the fixtures exercise them
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "LayerNorm" in n or "layer_norm" in n or n.endswith("bias"):
                p.add_(torch.randn(p.shape, generator=gen) * 0.05)
    m.eval()
    b = synth_batch(4, 12, 10, seed=11, variable=True, img_dim=IMG_DIM, vocab=TINY["vocab_size"],
                    min_txt=3, min_bb=2)
    kw = dict(input_ids=b["input_ids"], position_ids=b["position_ids"], img_feat=b["img_feat"],
              img_pos_feat=b["img_pos_feat"], attention_mask=b["attn_mask"], gather_index=b["gather_index"],
              output_all_encoded_layers=False)
    logits = m(**kw)
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([1.8]))
    loss = crit(logits.squeeze(1), b["labels"].float())
    loss.backward()
    all_layers = um(**{**kw, "output_all_encoded_layers": True})
    emb = um._compute_img_txt_embeddings(b["input_ids"], b["position_ids"], b["img_feat"], b["img_pos_feat"],
                                         b["gather_index"])
    out = {"sd." + k: v.detach().numpy() for k, v in m.state_dict().items()}
    out.update({"in." + k: (v.numpy() if torch.is_tensor(v) else np.array(v)) for k, v in b.items()})
    out["logits"] = logits.detach().numpy()
    out["loss"] = loss.detach().numpy()
    out["embedding_output"] = emb.detach().numpy()
    out["layer0_out"] = all_layers[0].detach().numpy()
    out["layer1_out"] = all_layers[1].detach().numpy()
    for n, p in m.named_parameters():
        if p.grad is not None:  # mask_embedding gets no grad without img_masks
            out["grad." + n] = p.grad.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_meme_uniter.npz"), **out)

    # ---- 3. OT (model/ot.py) on a ragged seeded batch
    torch.manual_seed(3)
    B, M, N, D = 3, 6, 9, 16
    txt = torch.randn(B, M, D); img = torch.randn(B, N, D)
    txt_pad = torch.zeros(B, M, dtype=torch.bool); img_pad = torch.zeros(B, N, dtype=torch.bool)
    txt_pad[1, 4:] = True; img_pad[1, 7:] = True; txt_pad[2, 2:] = True; img_pad[2, 5:] = True
    cost = rot.cost_matrix_cosine(txt, img)
    joint_pad = txt_pad.unsqueeze(-1) | img_pad.unsqueeze(-2)
    cost_m = cost.masked_fill(joint_pad, 0)
    txt_len = (M - txt_pad.sum(1)).float(); img_len = (N - img_pad.sum(1)).float()
    T = rot.ipot(cost_m, txt_len, txt_pad, img_len, img_pad, joint_pad, 0.5, 50, 1)
    dist = rot.optimal_transport_dist(txt, img, txt_pad, img_pad)
    np.savez_compressed(os.path.join(OUT, "ot.npz"), txt=txt.numpy(), img=img.numpy(),
                        txt_pad=txt_pad.numpy(), img_pad=img_pad.numpy(), cost=cost.numpy(),
                        T=T.numpy(), dist=dist.numpy())
    # ---- 4. tiny UniterForPretraining: per-task losses (model/pretrain.py), eval mode
    import model.pretrain as rpre
    from oracle.uniter_oracle import synth_pretrain_batch
    torch.manual_seed(1)
    pm = rpre.UniterForPretraining(cfg, IMG_DIM, LABEL_DIM)
    with torch.no_grad():
        for n, p in pm.named_parameters():
            if "LayerNorm" in n or "layer_norm" in n or n.endswith("bias") or ".net.2." in n:
                p.add_(torch.randn(p.shape, generator=gen) * 0.05)
    pm.eval()
    pb = synth_pretrain_batch(4, 12, 10, seed=21, img_dim=IMG_DIM, vocab=TINY["vocab_size"], label_dim=LABEL_DIM,
                              min_txt=3, min_bb=2)
    out = {"sd." + k: v.detach().numpy() for k, v in pm.state_dict().items()}
    for k, v in pb.items():
        if torch.is_tensor(v):
            out["in." + k] = v.numpy()
    for k, v in pb["ot_inputs"].items():
        out["ot." + k] = v.numpy() if torch.is_tensor(v) else np.array(v)
    with torch.no_grad():
        out["loss.mlm"] = pm(pb, "mlm").numpy()
        out["loss.mrfr"] = pm(pb, "mrfr").numpy()
        out["loss.itm"] = pm(pb, "itm").numpy()
        out["loss.mrc-kl"] = pm(pb, "mrc-kl").numpy()
        out["loss.mrc"] = pm(pb, "mrc").numpy()
        out["scores.mlm"] = pm(pb, "mlm", compute_loss=False).numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_pretrain.npz"), **out)
    print("golden vectors written to", os.path.abspath(OUT))
    for f in sorted(os.listdir(OUT)):
        print("  %-28s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
