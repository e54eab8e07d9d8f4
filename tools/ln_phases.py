"""clock64 phase stamps of the fused-LayerNorm GEMM epilogue vs the plain residual epilogue (C2 shapes)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from meme_challenge_b200 import _lib, ops

dev = "cuda"
L = _lib.lib()
M, H = 2624, 768
names = ["setup", "tma_issue_end", "first_landed", "mma_issued", "acc_ready", "epi_done", "exit"]
seed = torch.tensor([7], device=dev, dtype=torch.int64)
drop = _lib.dropout_t(seed, 3, 0.1)
for K in (768, 3072):
    a = torch.randn(M, K, device=dev).bfloat16()
    b = (torch.randn(H, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(H, device=dev)
    res = torch.randn(M, H, device=dev).bfloat16()
    gam, bet = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    for label, kw in (("plain", dict(epilogue=_lib.EPI_BIAS_DROP_RES)),
                      ("plain cl=2", dict(epilogue=_lib.EPI_BIAS_DROP_RES, cluster=2)),
                      ("fused LN", dict(epilogue=_lib.EPI_BIAS_DROP_RES_LN, ln=(gam, bet, 1e-12, mean, rstd)))):
        for _ in range(3):
            ops.gemm(a, b, bias=bias, res=res, drop=drop, **kw)
        stamps = torch.zeros(148 * 8, device=dev, dtype=torch.int64)
        L.b200u_gemm_debug_stamps(stamps.data_ptr())
        ops.gemm(a, b, bias=bias, res=res, drop=drop, **kw)
        torch.cuda.synchronize()
        L.b200u_gemm_debug_stamps(None)
        st = stamps.view(148, 8).cpu()
        st = st[st[:, 0] != 0]
        t0 = st[:, 0].min()
        rel = (st - st[:, :1]).float()
        print("K=%d %-10s ctas=%d entry spread=%d  whole grid: last exit - first entry = %d cycles" % (
            K, label, st.shape[0], int((st[:, 0] - t0).max()), int((st[:, 7] - t0).max())))
        print("     mean cycles since own entry: " + "  ".join("%s=%.0f" % (n, rel[:, i + 1].mean().item()) for i, n in enumerate(names)))
        print("     max  cycles since own entry: " + "  ".join("%s=%.0f" % (n, rel[:, i + 1].max().item()) for i, n in enumerate(names)))
