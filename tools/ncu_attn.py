"""ncu target: attention fwd + bwd (dropout on) at the C2 shape, two rounds (profile the second: -s 3 -c 3)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from meme_challenge_b200 import _lib, ops

dev = "cuda"
B, L, heads, H = 16, 164, 12, 768
M = B * L
seed = torch.tensor([7], device=dev, dtype=torch.int64)
qkv = (torch.randn(M, 3 * H, device=dev) * 0.5).bfloat16()
mask = torch.zeros(B, L, device=dev)
d = _lib.dropout_t(seed, 5, 0.1)
dctx = torch.randn(M, H, device=dev).bfloat16()
for _ in range(2):
    ctx, lse = ops.attention_fwd(qkv, mask, B, L, heads, H, drop=d)
    ops.attention_bwd(qkv, mask, ctx, dctx, lse, B, L, heads, H, drop=d)
    torch.cuda.synchronize()
print("done")
