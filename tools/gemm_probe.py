"""GPU bring-up probe for the tcgen05 GEMM: runs groups of cases in sub-processes (so a hang or
a sticky CUDA error in one group cannot take the rest down), compares against torch fp32 matmul
and prints one line per case. Usage: python tools/gemm_probe.py [group ...]
"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

GROUPS = ["kk", "kmn", "mnmn", "epi", "time"]


def _ref(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float() if b_mn else b.float().t()
    return A @ B


def _mk(M, N, K, a_mn, b_mn, dev):
    import torch
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    return a, b


def _report(tag, out, ref, tol):
    import torch
    err = (out.float() - ref).abs()
    scale = ref.abs().max().item() + 1e-6
    bad = (err > tol * scale)
    nbad = int(bad.sum().item())
    msg = "%-58s max_err=%.4g rel=%.4g bad=%d/%d" % (tag, err.max().item(), err.max().item() / scale,
                                                      nbad, err.numel())
    if nbad:
        idx = torch.nonzero(bad)[0].tolist()
        msg += " first_bad=%s got=%.4g want=%.4g" % (idx, out.float()[tuple(idx)].item(),
                                                     ref[tuple(idx)].item())
        rows = torch.nonzero(bad.any(1)).flatten()
        cols = torch.nonzero(bad.any(0)).flatten()
        msg += " bad_rows[%d..%d]x%d bad_cols[%d..%d]x%d" % (rows.min(), rows.max(), rows.numel(),
                                                              cols.min(), cols.max(), cols.numel())
    print(("PASS " if nbad == 0 else "FAIL ") + msg, flush=True)
    return nbad == 0


def run_group(group):
    import torch
    from meme_challenge_b200 import ops
    from meme_challenge_b200 import _lib
    dev = "cuda"
    torch.manual_seed(0)
    ok = True
    if group in ("kk", "kmn", "mnmn"):
        a_mn = group == "mnmn"
        b_mn = group in ("kmn", "mnmn")
        shapes = [(128, 128, 64), (128, 128, 256), (128, 256, 128), (256, 384, 192), (2624, 768, 768),
                  (2624, 2304, 768), (200, 136, 72), (768, 3072, 2624)]
        for (M, N, K) in shapes:
            for bn in (128, 256):
                a, b = _mk(M, N, K, a_mn, b_mn, dev)
                ref = _ref(a, b, a_mn, b_mn)
                out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, block_n=bn)
                torch.cuda.synchronize()
                ok &= _report("%s M=%d N=%d K=%d bn=%d store" % (group, M, N, K, bn), out, ref, 1e-2)
        # SIMT debug kernel against the same reference
        a, b = _mk(200, 136, 72, a_mn, b_mn, dev)
        out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, impl=1)
        ok &= _report("%s simt 200x136x72" % group, out, _ref(a, b, a_mn, b_mn), 1e-2)
    elif group == "epi":
        M, N, K = 2624, 768, 768
        a, b = _mk(M, N, K, False, False, dev)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev).bfloat16()
        ref = _ref(a, b, False, False)
        for impl in (0, 1):
            t = "impl%d " % impl
            out = ops.gemm(a, b, bias=bias, impl=impl)
            ok &= _report(t + "bias", out, ref + bias, 1e-2)
            u, g = ops.gemm(a, b, bias=bias, epilogue=_lib.EPI_BIAS_GELU, impl=impl)
            ok &= _report(t + "gelu.u", u, ref + bias, 1e-2)
            ok &= _report(t + "gelu.g", g, torch.nn.functional.gelu((ref + bias).bfloat16().float()), 1e-2)
            out = ops.gemm(a, b, bias=bias, res=res, epilogue=_lib.EPI_BIAS_DROP_RES, impl=impl)
            ok &= _report(t + "bias_res(p=0)", out, ref + bias + res.float(), 1e-2)
            out = ops.gemm(a, b, res=res, epilogue=_lib.EPI_ADD, impl=impl)
            ok &= _report(t + "add", out, ref + res.float(), 1e-2)
            out = ops.gemm(a, b, res=res, epilogue=_lib.EPI_DGELU, impl=impl)
            x = res.float().requires_grad_(True)
            torch.nn.functional.gelu(x).sum().backward()
            ok &= _report(t + "dgelu", out, ref * x.grad, 1e-2)
            acc = torch.ones(M, N, device=dev)
            ops.gemm(a, b, epilogue=_lib.EPI_ATOMIC_F32, out=acc, splits=3, impl=impl)
            ok &= _report(t + "atomic_f32 splits=3", acc, ref + 1.0, 2e-3)
            out = ops.gemm(a, b, bias=bias, epilogue=_lib.EPI_STORE_F32, impl=impl)
            ok &= _report(t + "store_f32", out, ref + bias, 2e-3)
        # dropout keep-rate and agreement between the two implementations (same counter RNG)
        seed = torch.tensor([1234567], device=dev, dtype=torch.int64)
        d = _lib.dropout_t(seed, 7, 0.1)
        zero_res = torch.zeros(M, N, device=dev).bfloat16()
        o0 = ops.gemm(a, b, bias=bias, res=zero_res, epilogue=_lib.EPI_BIAS_DROP_RES, drop=d, impl=0)
        o1 = ops.gemm(a, b, bias=bias, res=zero_res, epilogue=_lib.EPI_BIAS_DROP_RES, drop=d, impl=1)
        keep = (o0 != 0).float().mean().item()
        print("%s dropout keep-rate=%.4f (want 0.9), tc-vs-simt mask agreement=%.6f" % (
            "PASS" if abs(keep - 0.9) < 5e-3 else "FAIL", keep,
            ((o0 != 0) == (o1 != 0)).float().mean().item()), flush=True)
        ok &= _report("dropout values", o0, torch.where(o0 != 0, (ref + bias) / 0.9, torch.zeros_like(ref)), 1e-2)
    elif group == "time":
        cases = [("qkv fwd", 2624, 2304, 768, False, False, _lib.EPI_STORE),
                 ("attn-out fwd", 2624, 768, 768, False, False, _lib.EPI_STORE),
                 ("ffn1 fwd", 2624, 3072, 768, False, False, _lib.EPI_STORE),
                 ("ffn2 fwd", 2624, 768, 3072, False, False, _lib.EPI_STORE),
                 ("ffn1 dgrad", 2624, 768, 3072, False, True, _lib.EPI_STORE),
                 ("ffn2 dgrad", 2624, 3072, 768, False, True, _lib.EPI_STORE),
                 ("ffn1 wgrad", 3072, 768, 2624, True, True, _lib.EPI_ATOMIC_F32),
                 ("qkv wgrad", 2304, 768, 2624, True, True, _lib.EPI_ATOMIC_F32),
                 ("o wgrad", 768, 768, 2624, True, True, _lib.EPI_ATOMIC_F32),
                 ("big 8192^3", 8192, 8192, 8192, False, False, _lib.EPI_STORE)]
        for (name, M, N, K, a_mn, b_mn, epi) in cases:
            for bn in (128, 256):
                a, b = _mk(M, N, K, a_mn, b_mn, dev)
                out = torch.zeros(M, N, device=dev, dtype=torch.float32 if epi == _lib.EPI_ATOMIC_F32 else torch.bfloat16)
                for _ in range(3):
                    ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, epilogue=epi, out=out, block_n=bn)
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                iters = 20
                e0.record()
                for _ in range(iters):
                    ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, epilogue=epi, out=out, block_n=bn)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                # torch/cuBLAS for context
                A = a.t() if a_mn else a
                B = b if b_mn else b.t()
                for _ in range(3):
                    torch.matmul(A, B)
                e0.record()
                for _ in range(iters):
                    torch.matmul(A, B)
                e1.record()
                torch.cuda.synchronize()
                ms_t = e0.elapsed_time(e1) / iters
                fl = 2.0 * M * N * K
                print("TIME %-14s M=%d N=%d K=%d bn=%d: %.2f us %.1f TFLOP/s | cuBLAS %.2f us %.1f TFLOP/s" % (
                    name, M, N, K, bn, ms * 1e3, fl / ms / 1e9, ms_t * 1e3, fl / ms_t / 1e9), flush=True)
    return ok


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        ok = run_group(sys.argv[2])
        sys.exit(0 if ok else 1)
    groups = sys.argv[1:] or GROUPS
    for g in groups:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", g], timeout=300)
            print("GROUP %s rc=%d (%.1fs)" % (g, r.returncode, time.time() - t0), flush=True)
        except subprocess.TimeoutExpired:
            print("GROUP %s TIMEOUT (hang?)" % g, flush=True)
