"""clock64 phase stamps of the tcgen05 attention forward (C2 shape): per-CTA chain latency."""
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
for _ in range(3):
    ops.attention_fwd(qkv, mask, B, L, heads, H, drop=d)
n = B * heads * 2
stamps = torch.zeros(n * 8, device=dev, dtype=torch.int64)
L_.b200u_gemm_debug_stamps(stamps.data_ptr())
ops.attention_fwd(qkv, mask, B, L, heads, H, drop=d)
torch.cuda.synchronize()
L_.b200u_gemm_debug_stamps(None)
st = stamps.view(n, 8).cpu()
names = ["pdl", "qk_landed", "s_done", "pass1", "pass2", "o_ready", "exit"]
t0 = st[:, 0].min()
for label, sel in (("heavy (tile 0)", st[:B * heads]), ("light (tile 1)", st[B * heads:])):
    rel = (sel - sel[:, :1]).float()
    print(label, "entry: min %d max %d (cycles after first entry)" % (int((sel[:, 0] - t0).min()), int((sel[:, 0] - t0).max())))
    print("   mean since own entry: " + "  ".join("%s=%.0f" % (nm, rel[:, k + 1].mean().item()) for k, nm in enumerate(names)))
    print("   max  since own entry: " + "  ".join("%s=%.0f" % (nm, rel[:, k + 1].max().item()) for k, nm in enumerate(names)))
print("grid: last exit - first entry = %d cycles" % int((st[:, 7] - t0).max()))
