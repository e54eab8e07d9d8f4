"""Live roofline measurements of the hot path's kernels (bench.py `roofline` object, tools/).

Every kernel is timed as 20 back-to-back launches inside a captured CUDA graph (host launch gaps excluded)
with CUDA events on the launching stream, alternating between two operand sets so consecutive launches do
not hit the same lines. GEMMs are scored against the bf16 tensor-core peak with their algorithmic FLOPs
(2 M N K); the element-wise / row kernels against the HBM peak with their algorithmic bytes:
what the kernel must read and write once, not what it happens to move.
"""
import torch

from . import _lib, ops

REPS = 20


def _time_graph(fn, reps=REPS):
    fn(0)
    fn(1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i & 1)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps  # us per launch


def layer_gemm_shapes(M, H, I):
    """(name, M, N, K, a_mn, b_mn, epilogue) of every GEMM of one BertLayer forward + backward as
    b200u_bert_layer_{fwd,bwd} launches them (csrc/layer.cu)."""
    E = _lib
    return [("qkv_fwd", M, 3 * H, H, 0, 0, E.EPI_STORE),
            ("attn_out_fwd_ln", M, H, H, 0, 0, E.EPI_BIAS_DROP_RES_LN),
            ("ffn1_fwd_gelu", M, I, H, 0, 0, E.EPI_BIAS_GELU_DG),
            ("ffn2_fwd_ln", M, H, I, 0, 0, E.EPI_BIAS_DROP_RES_LN),
            ("ffn2_wgrad", H, I, M, 1, 1, E.EPI_ATOMIC_F32),
            ("ffn2_dgrad_mul", M, I, H, 0, 1, E.EPI_MUL),
            ("ffn1_wgrad", I, H, M, 1, 1, E.EPI_ATOMIC_F32),
            ("ffn1_dgrad", M, H, I, 0, 1, E.EPI_ADD),
            ("attn_out_wgrad", H, H, M, 1, 1, E.EPI_ATOMIC_F32),
            ("attn_out_dgrad", M, H, H, 0, 1, E.EPI_STORE),
            ("qkv_wgrad", 3 * H, H, M, 1, 1, E.EPI_ATOMIC_F32),
            ("qkv_dgrad", M, H, 3 * H, 0, 1, E.EPI_ADD)]


def time_gemm(dev, m, n, k, am, bm, ep):
    """In-graph time (us) of one GEMM shape with the epilogue inputs the layer gives it."""
    E = _lib
    sets = []
    seed = torch.tensor([7], device=dev, dtype=torch.int64)
    for _ in range(2):
        a = torch.randn((k, m) if am else (m, k), device=dev).bfloat16()
        b = (torch.randn((k, n) if bm else (n, k), device=dev) * 0.05).bfloat16()
        f32 = ep in (E.EPI_ATOMIC_F32, E.EPI_STORE_F32)
        kw = dict(a_mn=bool(am), b_mn=bool(bm), epilogue=ep,
                  out=torch.zeros(m, n, device=dev, dtype=torch.float32 if f32 else torch.bfloat16))
        if ep in E.EPI_HAS_BIAS:
            kw["bias"] = torch.randn(n, device=dev)
        if ep in E.EPI_HAS_RES:
            kw["res"] = torch.rand(m, n, device=dev).bfloat16()
        if ep in E.EPI_DUAL:
            kw["out2"] = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        if ep in (E.EPI_BIAS_DROP_RES, E.EPI_BIAS_DROP_RES_LN):
            kw["drop"] = _lib.dropout_t(seed, 3, 0.1)
        if ep == E.EPI_BIAS_DROP_RES_LN:
            kw["ln"] = (torch.ones(n, device=dev), torch.zeros(n, device=dev), 1e-12,
                        torch.empty(m, device=dev), torch.empty(m, device=dev))
        if ep == E.EPI_MUL:
            kw["colsum"] = torch.zeros(n, device=dev)
        sets.append((a, b, kw))
    return _time_graph(lambda i: ops.gemm(sets[i][0], sets[i][1], **sets[i][2]))


def gemm_family(dev, M, H, I, layers, passes, img_rows, img_dim=2048):
    """All GEMM shapes of one optimizer step. `passes` = forward/backward passes per step (2 micro-batches, or 1
    fused window with M doubled by the caller). Returns (ms per step, FLOPs per step, launches, {shape: us})."""
    E = _lib
    shapes = [(nm, passes * layers, m, n, k, am, bm, ep) for (nm, m, n, k, am, bm, ep) in layer_gemm_shapes(M, H, I)]
    shapes += [("img_linear_fwd", passes, img_rows, H, img_dim, 0, 0, E.EPI_STORE_F32),
               ("img_linear_wgrad", passes, H, img_dim, img_rows, 1, 1, E.EPI_ATOMIC_F32)]
    tot_ms = tot_fl = 0.0
    launches, detail = 0, {}
    for (nm, count, m, n, k, am, bm, ep) in shapes:
        us = time_gemm(dev, m, n, k, am, bm, ep)
        detail[nm] = {"us": round(us, 2), "tflops": round(2.0 * m * n * k / us / 1e6, 1)}
        tot_ms += count * us * 1e-3
        tot_fl += count * 2.0 * m * n * k
        launches += count
    return tot_ms, tot_fl, launches, detail


def hbm_kernels(dev, B, T, R, L, H, heads, n_params, hbm_gbs):
    """HBM-bound kernels of the path: algorithmic bytes / in-graph time vs the measured copy peak.
    Returns {kernel: {us, bytes, gbs, frac}}."""
    from . import functional as F_  # noqa: F401  (keeps the package import order)
    out = {}
    M = B * L
    seed = torch.tensor([7], device=dev, dtype=torch.int64)
    d = _lib.dropout_t(seed, 5, 0.1)

    def rec(name, us, nbytes):
        out[name] = {"us": round(us, 2), "bytes": int(nbytes), "gbs": round(nbytes / us / 1e3, 1),
                     "frac": round(nbytes / us / 1e3 / hbm_gbs, 4)}

    # LayerNorm backward (dy, y in; dx, dz out) as the layer calls it
    x = [torch.randn(M, H, device=dev).bfloat16() for _ in range(2)]
    dy = [torch.randn(M, H, device=dev).bfloat16() for _ in range(2)]
    gam, bet = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    _, mean, rstd = ops.layernorm_fwd(x[0], gam, bet, 1e-12)
    dg, db_, dbias = (torch.zeros(H, device=dev) for _ in range(3))
    rec("layernorm_bwd", _time_graph(lambda i: ops.layernorm_bwd(dy[i], x[i], mean, rstd, gam, dg, db_, dz=True,
                                                                 dbias=dbias, drop=_lib.dropout_t(seed, 3, 0.1))),
        4 * M * H * 2)
    rec("layernorm_fwd", _time_graph(lambda i: ops.layernorm_fwd(x[i], gam, bet, 1e-12)), 2 * M * H * 2)
    # attention forward / backward: Q, K, V in + ctx out ; backward adds dO in and dQ, dK, dV out
    qkv = [(torch.randn(M, 3 * H, device=dev) * 0.5).bfloat16() for _ in range(2)]
    mask = torch.zeros(B, L, device=dev)
    ctx, lse = ops.attention_fwd(qkv[0], mask, B, L, heads, H, drop=d)
    dctx = torch.randn(M, H, device=dev).bfloat16()
    rec("attention_fwd", _time_graph(lambda i: ops.attention_fwd(qkv[i], mask, B, L, heads, H, drop=d)), 4 * M * H * 2)
    dbq = torch.zeros(3 * H, device=dev)
    rec("attention_bwd", _time_graph(lambda i: ops.attention_bwd(qkv[i], mask, ctx, dctx, lse, B, L, heads, H, drop=d,
                                                                 dbias_qkv=dbq)), 9 * M * H * 2)
    # K0 text embeddings: word + position rows in (fp32 tables), LayerNorm'd bf16 row out (+ the type row, negligible);
    # SURVEY 8(d) counts 3 reads + 1 write of T*H bf16 = 0.39 MB per meme
    import ctypes as C
    P = _lib.ptr
    V, PMAX = 28996, 512
    word = torch.randn(V, H, device=dev) * 0.02
    post = torch.randn(PMAX, H, device=dev) * 0.02
    typ = torch.randn(2, H, device=dev) * 0.02
    ids = [torch.randint(1000, V, (B, T), device=dev) for _ in range(2)]
    pos_ids = torch.arange(T, device=dev).unsqueeze(0).repeat(B, 1).contiguous()
    t_out = torch.empty(B * T, H, device=dev, dtype=torch.bfloat16)
    t_sum = torch.empty(B * T, H, device=dev)
    t_mean, t_rstd = torch.empty(B * T, device=dev), torch.empty(B * T, device=dev)
    rec("txt_embed_fwd", _time_graph(lambda i: ops._call(
        "b200u_txt_embed_fwd", P(ids[i]), P(pos_ids), T, None, P(word), P(post), P(typ), P(gam), P(bet), P(t_out),
        P(t_sum), P(t_mean), P(t_rstd), B, T, H, V, PMAX, 2, 1e-12, C.byref(d))), 4 * B * T * H * 2)
    # K2 image embeddings after img_linear: fp32 projection row + 7-d box in, LayerNorm'd bf16 row out
    n_img = B * R
    a_in = [torch.randn(n_img, H, device=dev) for _ in range(2)]
    pos7 = torch.rand(n_img, 7, device=dev)
    Wp, bp = torch.randn(H, 7, device=dev) * 0.1, torch.zeros(H, device=dev)
    i_out = torch.empty(n_img, H, device=dev, dtype=torch.bfloat16)
    i_p, i_s = torch.empty(n_img, H, device=dev), torch.empty(n_img, H, device=dev)
    i_stats = torch.empty(6, n_img, device=dev)
    rec("img_embed_fwd", _time_graph(lambda i: ops._call(
        "b200u_img_embed_fwd", P(a_in[i]), P(pos7), P(Wp), P(bp), None, P(typ), P(gam), P(bet), P(gam), P(bet), P(gam),
        P(bet), P(i_out), P(i_p), P(i_s), P(i_stats), n_img, H, 2, 1e-12, C.byref(d))), n_img * (H * 6 + 28))
    # gather_index concat: read txt + img rows once, write the joint sequence
    txt = [torch.randn(B, T, H, device=dev).bfloat16() for _ in range(2)]
    img = [torch.randn(B, R, H, device=dev).bfloat16() for _ in range(2)]
    gi = torch.arange(L, device=dev).unsqueeze(0).repeat(B, 1).contiguous()
    rec("gather_rows", _time_graph(lambda i: ops.gather_rows(txt[i], img[i], gi)), 2 * M * H * 2)
    # fused Adam (+ clip + shadow + zero_grad): 16 B read + 18 B written per parameter
    n = (n_params + 1023) // 1024 * 1024
    p, g, m_, v = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
    v.abs_()
    sh = torch.empty(n, device=dev, dtype=torch.bfloat16)
    run_start = torch.tensor([0, n], device=dev, dtype=torch.int64)
    run_wd = torch.tensor([1e-3], device=dev)
    chunk_run = torch.zeros(n // 1024, device=dev, dtype=torch.int32)
    coef, lr = torch.ones(1, device=dev), torch.tensor([3e-5], device=dev)
    step = torch.ones(1, device=dev, dtype=torch.int64)
    rec("adam_step", _time_graph(lambda i: ops._call("b200u_adam_step", P(p), P(g), P(m_), P(v), P(sh), C.c_size_t(n),
                                                     P(run_start), P(run_wd), P(chunk_run), 1, P(coef), P(lr), P(step),
                                                     0.9, 0.999, 1e-8, 1, None, C.c_size_t(0), C.c_size_t(0)), reps=4),
        34 * n)
    return out


def ot_kernels(dev, B=16, M=64, N=100, D=768):
    """IPOT / OT-distance launch times (latency-bound: one CTA per sample) vs the launch floor."""
    from .model import ot
    x = [torch.randn(B, M, D, device=dev) for _ in range(2)]
    y = [torch.randn(B, N, D, device=dev) for _ in range(2)]
    xp = torch.zeros(B, M, dtype=torch.bool, device=dev)
    yp = torch.zeros(B, N, dtype=torch.bool, device=dev)
    us = _time_graph(lambda i: ot.optimal_transport_dist(x[i], y[i], xp, yp), reps=10)
    nul = torch.zeros(1, device=dev, dtype=torch.int64)
    floor = _time_graph(lambda i: ops.counter_add(nul, 0))
    return {"optimal_transport_dist_us": round(us, 2), "launch_floor_us": round(floor, 2),
            "what": "cosine cost + 50 IPOT iterations (one launch) + trace distance for %d samples of %d x %d" % (B, M, N)}
