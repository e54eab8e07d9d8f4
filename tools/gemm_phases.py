"""Per-CTA phase timing of the tcgen05 GEMM (clock64 stamps) for the C2 shapes."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from meme_challenge_b200 import _lib, ops

dev = "cuda"
L = _lib.lib()
cases = [("qkv fwd", 2624, 2304, 768, False, False, _lib.EPI_STORE, 0),
         ("o fwd", 2624, 768, 768, False, False, _lib.EPI_BIAS_DROP_RES, 0),
         ("ffn1 fwd", 2624, 3072, 768, False, False, _lib.EPI_BIAS_GELU, 0),
         ("ffn2 fwd", 2624, 768, 3072, False, False, _lib.EPI_BIAS_DROP_RES, 0),
         ("ffn2 dgrad", 2624, 3072, 768, False, True, _lib.EPI_DGELU, 0),
         ("ffn1 dgrad", 2624, 768, 3072, False, True, _lib.EPI_ADD, 0),
         ("ffn1 wgrad", 3072, 768, 2624, True, True, _lib.EPI_ATOMIC_F32, 0),
         ("o wgrad", 768, 768, 2624, True, True, _lib.EPI_ATOMIC_F32, 0)]
names = ["setup", "tma_issue_end", "first_landed", "mma_issued", "acc_ready", "epi_done", "exit"]
for (name, M, N, K, a_mn, b_mn, epi, _) in cases:
    for bn, cl, mode in ((128, 1, 0), (128, 1, 2), (128, 2, 0), (256, 1, 0), (256, 1, 2), (256, 2, 0)):
        a = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
        b = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev).bfloat16()
        out = torch.zeros(M, N, device=dev, dtype=torch.float32 if epi == _lib.EPI_ATOMIC_F32 else torch.bfloat16)
        kw = dict(a_mn=a_mn, b_mn=b_mn, epilogue=epi, out=out, block_n=bn, cluster=cl)
        if epi in (_lib.EPI_BIAS_DROP_RES, _lib.EPI_BIAS_GELU): kw["bias"] = bias
        if epi in (_lib.EPI_BIAS_DROP_RES, _lib.EPI_ADD, _lib.EPI_DGELU): kw["res"] = res
        for _ in range(3):
            ops.gemm(a, b, **kw)
        stamps = torch.zeros(148 * 8, device=dev, dtype=torch.int64)
        L.b200u_gemm_debug_stamps(stamps.data_ptr() + mode)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, b, **kw); e1.record()
        torch.cuda.synchronize()
        L.b200u_gemm_debug_stamps(None)
        st = stamps.view(148, 8).cpu()
        used = st[:, 0] != 0
        st = st[used]
        t0 = st[:, 0].min()
        rel = (st - st[:, :1]).float()   # per-CTA relative to own entry
        start_spread = (st[:, 0] - t0).float()
        end = (st[:, 7] - t0).float()
        nkb = (K + 63) // 64
        print("%-11s mode=%d bn=%d cl=%d M=%d N=%d K=%d ctas=%d  cyc/kblock(first tile)=%.0f  exit mean=%.0f max=%.0f" % (
            name, mode, bn, cl, M, N, K, int(used.sum()), (rel[:, 5].mean().item() - rel[:, 3].mean().item()) / max(1, min(nkb, 9999)) if epi != _lib.EPI_ATOMIC_F32 else float("nan"),
            rel[:, 7].mean().item(), rel[:, 7].max().item()))
        print("     per-CTA mean cycles since entry: " + "  ".join("%s=%.0f" % (n, rel[:, i + 1].mean().item()) for i, n in enumerate(names)))
        print("     per-CTA max  cycles since entry: " + "  ".join("%s=%.0f" % (n, rel[:, i + 1].max().item()) for i, n in enumerate(names)))
