"""ncu target: every GEMM shape of one BertLayer (forward + backward, the epilogues csrc/layer.cu uses) launched
three rounds in a row at M = GB_M rows (default 5248 = fused window). Capture round 3:
    ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 24 -c 12 -o prof python tools/ncu_target.py
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from meme_challenge_b200 import _lib, ops, roofline

dev = "cuda"
E = _lib
M, H, I = int(os.environ.get("GB_M", "5248")), 768, 3072
seed = torch.tensor([7], device=dev, dtype=torch.int64)
sets = []
for (nm, m, n, k, am, bm, ep) in roofline.layer_gemm_shapes(M, H, I):
    a = torch.randn((k, m) if am else (m, k), device=dev).bfloat16()
    b = (torch.randn((k, n) if bm else (n, k), device=dev) * 0.05).bfloat16()
    f32 = ep in (E.EPI_ATOMIC_F32, E.EPI_STORE_F32)
    kw = dict(a_mn=bool(am), b_mn=bool(bm), epilogue=ep,
              out=torch.zeros(m, n, device=dev, dtype=torch.float32 if f32 else torch.bfloat16))
    if ep in E.EPI_HAS_BIAS: kw["bias"] = torch.randn(n, device=dev)
    if ep in E.EPI_HAS_RES: kw["res"] = torch.rand(m, n, device=dev).bfloat16()
    if ep in E.EPI_DUAL: kw["out2"] = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    if ep in (E.EPI_BIAS_DROP_RES, E.EPI_BIAS_DROP_RES_LN): kw["drop"] = _lib.dropout_t(seed, 3, 0.1)
    if ep == E.EPI_BIAS_DROP_RES_LN:
        kw["ln"] = (torch.ones(n, device=dev), torch.zeros(n, device=dev), 1e-12, torch.empty(m, device=dev), torch.empty(m, device=dev))
    if ep == E.EPI_MUL: kw["colsum"] = torch.zeros(n, device=dev)
    sets.append((nm, a, b, kw))
for r in range(3):
    for nm, a, b, kw in sets:
        ops.gemm(a, b, **kw)
    torch.cuda.synchronize()
print("done", [s[0] for s in sets])
