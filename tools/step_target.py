"""ncu / sanitizer target: ONE eager (no CUDA graph) fused-window optimizer step of UNITER-base at the C2 shape
(after one warm-up step), so every kernel of the path appears once per layer:
    ncu --set full --clock-control none -k regex:'txt_embed|img_embed|gather_rows|layernorm|attn_|adam|sumsq' ...
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from meme_challenge_b200.data.synthetic import synth_batch
from meme_challenge_b200.model.meme_uniter import MemeUniter
from meme_challenge_b200.model.model import UniterConfig, UniterModel
from meme_challenge_b200.train import TrainStep

layers = int(os.environ.get("ST_LAYERS", "2"))
cfg = dict(vocab_size=28996, hidden_size=768, num_hidden_layers=layers, num_attention_heads=12, intermediate_size=3072,
           hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=512,
           type_vocab_size=2, initializer_range=0.02)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = MemeUniter(UniterModel(UniterConfig.from_dict(cfg), 2048), 768, 1).to(dev).train()
ts = TrainStep(m, gradient_accumulation=2, fuse_window=True)
bs = []
for i in range(2):
    b = synth_batch(16, 64, 100, seed=1234 + i)
    b = {k: v.to(dev) for k, v in b.items() if torch.is_tensor(v)}
    b["labels"] = b["labels"].float()
    bs.append(b)
for _ in range(int(os.environ.get("ST_STEPS", "2"))):
    ts.step(bs)
torch.cuda.synchronize()
print("done")
