"""Launch a fixed list of GEMM shapes three rounds in a row (ncu target: -s 2*len -c len captures round 3)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from meme_challenge_b200 import _lib, ops

dev = "cuda"
E = _lib
M, H, I = 2624, 768, 3072
ALL = {"ffn1_fwd": (M, I, H, 0, 0, E.EPI_BIAS_GELU, 256, 0), "ffn2_dgrad": (M, I, H, 0, 1, E.EPI_DGELU, 256, 0),
       "ffn2_wgrad": (H, I, M, 1, 1, E.EPI_ATOMIC_F32, 256, 2), "qkv_fwd": (M, 3 * H, H, 0, 0, E.EPI_STORE, 256, 0),
       "ffn2_fwd": (M, H, I, 0, 0, E.EPI_BIAS_DROP_RES, 128, 0), "attn_out_fwd": (M, H, H, 0, 0, E.EPI_BIAS_DROP_RES, 128, 0)}
names = sys.argv[1:] or list(ALL)
sets = []
for nme in names:
    m, n, k, am, bm, ep, bn, sp = ALL[nme]
    a = torch.randn((k, m) if am else (m, k), device=dev).bfloat16()
    b = torch.randn((k, n) if bm else (n, k), device=dev).bfloat16()
    f32 = ep in (E.EPI_ATOMIC_F32, E.EPI_STORE_F32)
    kw = dict(a_mn=bool(am), b_mn=bool(bm), epilogue=ep, block_n=bn, splits=sp,
              out=torch.zeros(m, n, device=dev, dtype=torch.float32 if f32 else torch.bfloat16))
    if ep in (E.EPI_STORE, E.EPI_BIAS_DROP_RES, E.EPI_BIAS_GELU):
        kw["bias"] = torch.randn(n, device=dev)
    if ep in (E.EPI_BIAS_DROP_RES, E.EPI_ADD, E.EPI_DGELU):
        kw["res"] = torch.randn(m, n, device=dev).bfloat16()
    if ep == E.EPI_BIAS_GELU:
        kw["out2"] = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    sets.append((a, b, kw))
for r in range(3):
    for a, b, kw in sets:
        ops.gemm(a, b, **kw)
    torch.cuda.synchronize()
print("done", names)
