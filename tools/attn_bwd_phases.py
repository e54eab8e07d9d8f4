"""clock64 phase stamps of the tcgen05 attention backward (C2 shape)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from meme_challenge_b200 import _lib, ops
dev = "cuda"
L_ = _lib.lib()
B, L, heads, H = int(os.environ.get("KB_B", "16")), 164, 12, 768
M = B * L
seed = torch.tensor([7], device=dev, dtype=torch.int64)
qkv = (torch.randn(M, 3 * H, device=dev) * 0.5).bfloat16()
mask = torch.zeros(B, L, device=dev)
d = _lib.dropout_t(seed, 5, 0.1)
ctx, lse = ops.attention_fwd(qkv, mask, B, L, heads, H, drop=d)
dctx = torch.randn(M, H, device=dev).bfloat16()
dbias = torch.zeros(3 * H, device=dev)
for _ in range(3):
    ops.attention_bwd(qkv, mask, ctx, dctx, lse, B, L, heads, H, drop=d, dbias_qkv=dbias)
n = B * heads
stamps = torch.zeros(n * 16, device=dev, dtype=torch.int64)
L_.b200u_gemm_debug_stamps(stamps.data_ptr())
ops.attention_bwd(qkv, mask, ctx, dctx, lse, B, L, heads, H, drop=d, dbias_qkv=dbias)
torch.cuda.synchronize()
L_.b200u_gemm_debug_stamps(None)
st = stamps.view(n, 16).cpu()
names = {1: "loaded", 2: "prologue", 3: "s0_ready", 4: "ew0_done", 5: "o0_ready", 6: "rd0_done", 7: "s1_ready", 8: "ew1_done",
         9: "o1_ready", 10: "rd1_done", 11: "dq_done", 12: "exit"}
rel = (st - st[:, :1]).float()
t0 = st[:, 0].min()
print("CTAs %d; entry spread %d cycles; grid span %d cycles" % (n, int((st[:, 0] - t0).max()), int((st[:, 12] - t0).max())))
print("  mean since own entry: " + "  ".join("%s=%.0f" % (names[k], rel[:, k].mean().item()) for k in sorted(names)))
print("  max  since own entry: " + "  ".join("%s=%.0f" % (names[k], rel[:, k].max().item()) for k in sorted(names)))
