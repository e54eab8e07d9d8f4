import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from oracle import uniter_oracle as O
from oracle.make_golden import IMG_DIM, TINY
from meme_challenge_b200.model.meme_uniter import MemeUniter
from meme_challenge_b200.model.model import UniterConfig, UniterModel
DEV = "cuda"

def build(cfgd):
    torch.manual_seed(0)
    cfg = UniterConfig.from_dict(cfgd)
    return MemeUniter(UniterModel(cfg, IMG_DIM), cfg.hidden_size, 1).to(DEV)

def kw(b):
    return dict(input_ids=b["input_ids"].to(DEV), position_ids=b["position_ids"].to(DEV),
                img_feat=b["img_feat"].to(DEV), img_pos_feat=b["img_pos_feat"].to(DEV),
                attention_mask=b["attn_mask"].to(DEV), gather_index=b["gather_index"].to(DEV),
                output_all_encoded_layers=False)

b = O.synth_batch(4, 12, 10, seed=3, img_dim=IMG_DIM, vocab=TINY["vocab_size"], min_txt=2, min_bb=2)
names = ["uniter_model.encoder.layer.1.output.dense.weight", "uniter_model.encoder.layer.0.attention.self.query.weight",
         "uniter_model.encoder.layer.0.attention.self.query.bias", "uniter_model.encoder.layer.1.output.LayerNorm.weight",
         "uniter_model.pooler.dense.weight", "uniter_model.embeddings.word_embeddings.weight", "linear.weight",
         "uniter_model.img_embeddings.img_linear.weight", "uniter_model.encoder.layer.1.intermediate.dense.bias"]
for impl in (0, 1):
    m = build(TINY).eval(); m.uniter_model.gemm_impl = impl
    ps = dict(m.named_parameters())
    m(**kw(b)).sum().backward()
    g1 = {n: ps[n].grad.clone() for n in names}
    ptr1 = {n: ps[n].grad.data_ptr() for n in names}
    m(**kw(b)).sum().backward()
    for n in names:
        g2 = ps[n].grad
        r = (g2.flatten() @ g1[n].flatten() / (g1[n].flatten() @ g1[n].flatten() + 1e-30)).item()
        print("impl%d eval accum %-62s ratio=%.4f |g1|=%.3e |g2|=%.3e sameptr=%s" % (impl, n, r, g1[n].norm().item(), g2.norm().item(), ptr1[n] == g2.data_ptr()))
for (ph, pa) in ((0.1, 0.1), (0.1, 0.0), (0.0, 0.1)):
    c = dict(TINY); c["hidden_dropout_prob"] = ph; c["attention_probs_dropout_prob"] = pa
    m = build(c).train()
    ps = dict(m.named_parameters())
    l1 = m(**kw(b))
    l1.sum().backward()
    print("train ph=%.1f pa=%.1f single:" % (ph, pa), {n.split("uniter_model.")[-1][-40:]: "%.3e" % ps[n].grad.norm().item() for n in names})
    m2 = build(c).train(); ps2 = dict(m2.named_parameters())
    a = m2(**kw(b)); bb = m2(**kw(b)); bb.sum().backward()
    print("train ph=%.1f pa=%.1f two-fwd:" % (ph, pa), {n.split("uniter_model.")[-1][-40:]: "%.3e" % ps2[n].grad.norm().item() for n in names})
