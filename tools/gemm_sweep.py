"""In-graph timing of every GEMM shape of the C2 step at both tile widths, plus per-CTA phase stamps
for the two GELU epilogues and the kernel-to-kernel spacing floor of a captured graph."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from meme_challenge_b200 import _lib, ops

dev = "cuda"
E = _lib
L = _lib.lib()
M, H, I = 2624, 768, 3072
shapes = [("qkv_fwd", M, 3 * H, H, 0, 0, E.EPI_STORE), ("attn_out_fwd", M, H, H, 0, 0, E.EPI_BIAS_DROP_RES),
          ("ffn1_fwd", M, I, H, 0, 0, E.EPI_BIAS_GELU), ("ffn1_fwd_nogelu", M, I, H, 0, 0, E.EPI_STORE),
          ("ffn2_fwd", M, H, I, 0, 0, E.EPI_BIAS_DROP_RES),
          ("ffn2_wgrad", H, I, M, 1, 1, E.EPI_ATOMIC_F32), ("ffn2_dgrad", M, I, H, 0, 1, E.EPI_DGELU),
          ("ffn2_dgrad_add", M, I, H, 0, 1, E.EPI_ADD), ("ffn2_dgrad_store", M, I, H, 0, 1, E.EPI_STORE),
          ("ffn1_wgrad", I, H, M, 1, 1, E.EPI_ATOMIC_F32), ("ffn1_dgrad", M, H, I, 0, 1, E.EPI_ADD),
          ("attn_out_wgrad", H, H, M, 1, 1, E.EPI_ATOMIC_F32), ("attn_out_dgrad", M, H, H, 0, 1, E.EPI_STORE),
          ("qkv_wgrad", 3 * H, H, M, 1, 1, E.EPI_ATOMIC_F32), ("qkv_dgrad", M, H, 3 * H, 0, 1, E.EPI_ADD)]


def mk(m, n, k, am, bm, ep):
    a = torch.randn((k, m) if am else (m, k), device=dev).bfloat16()
    b = torch.randn((k, n) if bm else (n, k), device=dev).bfloat16()
    f32 = ep in (E.EPI_ATOMIC_F32, E.EPI_STORE_F32)
    kw = dict(a_mn=bool(am), b_mn=bool(bm), epilogue=ep,
              out=torch.zeros(m, n, device=dev, dtype=torch.float32 if f32 else torch.bfloat16))
    if ep in (E.EPI_STORE, E.EPI_BIAS_DROP_RES, E.EPI_BIAS_GELU, E.EPI_STORE_F32):
        kw["bias"] = torch.randn(n, device=dev)
    if ep in (E.EPI_BIAS_DROP_RES, E.EPI_ADD, E.EPI_DGELU):
        kw["res"] = torch.randn(m, n, device=dev).bfloat16()
    if ep == E.EPI_BIAS_GELU:
        kw["out2"] = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    return a, b, kw


def graph_time(fn, reps=20):
    fn(0); fn(1); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, 1e3 * e0.elapsed_time(e1) / reps)
    return best


def sweep(tag):
    cnt = torch.zeros(1, device=dev, dtype=torch.int64)
    print("[%s] graph spacing floor (counter_add x20): %.2f us per launch" % (tag, graph_time(lambda i: ops.counter_add(cnt, 1))), flush=True)
    # fixed cost of one GEMM launch: a single 128x128x64 tile
    ta, tb, tkw = mk(128, 128, 64, 0, 0, E.EPI_STORE)
    print("[%s] one-tile GEMM (128x128x64) x20: %.2f us per launch" % (tag, graph_time(lambda i: ops.gemm(ta, tb, block_n=128, **tkw))), flush=True)
    for (name, m, n, k, am, bm, ep) in shapes:
        sets = [mk(m, n, k, am, bm, ep) for _ in range(2)]
        row = []
        combos = [(128, 0), (256, 0)]
        if ep == E.EPI_ATOMIC_F32:
            combos = [(bn, sp) for bn in (128, 256) for sp in (1, 2, 3, 4, 6, 8)]
        for bn, sp in combos:
            def fn(i, bn=bn, sp=sp):
                a, b, kw = sets[i % 2]
                ops.gemm(a, b, block_n=bn, splits=sp, **kw)
            try:
                us = graph_time(fn)
                row.append("bn%d%s %6.2f us %4.0f TF/s" % (bn, ("/s%d" % sp) if sp else "", us, 2.0 * m * n * k / us / 1e6))
            except Exception as e:  # noqa: BLE001
                row.append("bn%d failed: %s" % (bn, str(e)[:60]))
        print("[%s] %-18s M=%4d N=%4d K=%4d  %s" % (tag, name, m, n, k, "  ".join(row)), flush=True)


if os.environ.get("SWEEP_PDL0"):
    L.b200u_set_pdl(0)
    sweep("pdl=0")
L.b200u_set_pdl(1)
sweep("pdl=1")
if os.environ.get("SWEEP_NO_STAMPS"):
    sys.exit(0)

# phase stamps for the expensive epilogues
names = ["setup", "tma_issue_end", "first_landed", "mma_issued", "acc_ready", "epi_done", "exit"]
for (name, m, n, k, am, bm, ep) in [s for s in shapes if s[0] in ("ffn1_fwd", "ffn1_fwd_nogelu", "ffn2_dgrad", "ffn2_dgrad_add", "ffn2_dgrad_store", "qkv_fwd", "ffn2_fwd", "ffn2_wgrad", "attn_out_wgrad")]:
    for bn in (128, 256):
        a, b, kw = mk(m, n, k, am, bm, ep)
        for _ in range(3):
            ops.gemm(a, b, block_n=bn, **kw)
        stamps = torch.zeros(148 * 8, device=dev, dtype=torch.int64)
        L.b200u_gemm_debug_stamps(stamps.data_ptr())
        ops.gemm(a, b, block_n=bn, **kw)
        torch.cuda.synchronize()
        L.b200u_gemm_debug_stamps(None)
        st = stamps.view(148, 8).cpu()
        st = st[st[:, 0] != 0]
        t0 = st[:, 0].min()
        rel = (st - st[:, :1]).float()
        print("%-18s bn=%d ctas=%d | mean: %s | max: %s" % (
            name, bn, st.shape[0],
            " ".join("%s=%.0f" % (nm, rel[:, i + 1].mean().item()) for i, nm in enumerate(names)),
            " ".join("%s=%.0f" % (nm, rel[:, i + 1].max().item()) for i, nm in enumerate(names))), flush=True)
