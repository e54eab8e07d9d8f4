"""clock64 phase stamps of selected GEMM shapes (per-CTA, cycles since own entry)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from meme_challenge_b200 import _lib, ops, roofline
dev = "cuda"; L = _lib.lib(); E = _lib
M, H, I = int(os.environ.get("GB_M", "5248")), 768, 3072
names = ["setup", "tma_issue_end", "first_landed", "mma_issued", "acc_ready", "epi_done", "exit"]
seed = torch.tensor([7], device=dev, dtype=torch.int64)
GLOBAL = "--global" in sys.argv   # stamps from the device-wide ns clock: entry skew and true grid span
sys.argv = [a for a in sys.argv if a != "--global"]
want = sys.argv[1:] or ["qkv_fwd", "ffn1_fwd_gelu", "ffn2_fwd_ln", "ffn2_dgrad_mul", "ffn1_dgrad", "attn_out_fwd_ln"]
shapes = {s[0]: s for s in roofline.layer_gemm_shapes(M, H, I)}
shapes["ffn1_fwd_store"] = ("ffn1_fwd_store", M, I, H, 0, 0, E.EPI_STORE)
for nm in want:
    _, m, n, k, am, bm, ep = shapes[nm]
    a = torch.randn((k, m) if am else (m, k), device=dev).bfloat16()
    b = (torch.randn((k, n) if bm else (n, k), device=dev) * 0.05).bfloat16()
    kw = dict(a_mn=bool(am), b_mn=bool(bm), epilogue=ep, out=torch.zeros(m, n, device=dev, dtype=torch.bfloat16))
    if ep in E.EPI_HAS_BIAS: kw["bias"] = torch.randn(n, device=dev)
    if ep in E.EPI_HAS_RES: kw["res"] = torch.rand(m, n, device=dev).bfloat16()
    if ep in E.EPI_DUAL: kw["out2"] = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    if ep in (E.EPI_BIAS_DROP_RES, E.EPI_BIAS_DROP_RES_LN): kw["drop"] = _lib.dropout_t(seed, 3, 0.1)
    if ep == E.EPI_BIAS_DROP_RES_LN:
        kw["ln"] = (torch.ones(n, device=dev), torch.zeros(n, device=dev), 1e-12, torch.empty(m, device=dev), torch.empty(m, device=dev))
    if ep == E.EPI_MUL: kw["colsum"] = torch.zeros(n, device=dev)
    for _ in range(3):
        ops.gemm(a, b, **kw)
    stamps = torch.zeros(148 * 8, device=dev, dtype=torch.int64)
    L.b200u_gemm_debug_stamps(stamps.data_ptr() | (3 if GLOBAL else 0))
    ops.gemm(a, b, **kw)
    torch.cuda.synchronize()
    L.b200u_gemm_debug_stamps(None)
    st = stamps.view(148, 8).cpu(); st = st[st[:, 0] != 0]
    if GLOBAL:
        t0 = st[:, 0].min()
        print("%-16s ctas=%3d (ns) entry skew=%d  first exit=%d  last exit=%d | mean lifetime=%.0f max=%d" % (
            nm, st.shape[0], int(st[:, 0].max() - t0), int(st[:, 7].min() - t0), int(st[:, 7].max() - t0),
            (st[:, 7] - st[:, 0]).float().mean().item(), int((st[:, 7] - st[:, 0]).max())))
        ent = (st[:, 0] - t0).sort().values
        print("%-16s entry times (ns) of CTAs by rank: " % "" + " ".join(str(int(ent[i])) for i in range(0, ent.numel(), max(1, ent.numel() // 12))))
        continue
    rel = (st - st[:, :1]).float()
    t0 = st[:, 0].min()
    print("%-16s ctas=%3d grid span=%6d cycles | mean: " % (nm, st.shape[0], int((st[:, 7] - t0).max())) +
          "  ".join("%s=%.0f" % (x, rel[:, i + 1].mean().item()) for i, x in enumerate(names)))
    print("%-16s %28s | max : " % ("", "") + "  ".join("%s=%.0f" % (x, rel[:, i + 1].max().item()) for i, x in enumerate(names)))
