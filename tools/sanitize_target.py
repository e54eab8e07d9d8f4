"""compute-sanitizer target: one small launch of every kernel family (all GEMM epilogues and operand
layouts, attention fwd / fused bwd / split bwd, LayerNorm fwd / bwd, embeddings, gather, heads, OT,
optimizer). Shapes are small so memcheck / racecheck / synccheck finish in a minute each:

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py
    compute-sanitizer --tool synccheck python tools/sanitize_target.py
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from meme_challenge_b200 import _lib, ops  # noqa: E402

dev = "cuda"
E = _lib
only = set(sys.argv[1:])


def want(name):
    return not only or name in only


seed = torch.tensor([7], device=dev, dtype=torch.int64)
torch.manual_seed(0)

if want("gemm"):
    M, N, K = 304, 320, 192   # ragged in M (3 tiles), 2-3 n-tiles, 3 k-blocks
    for (am, bm) in ((0, 0), (0, 1), (1, 1), (1, 0)):
        a = torch.randn((K, M) if am else (M, K), device=dev).bfloat16()
        b = torch.randn((K, N) if bm else (N, K), device=dev).bfloat16()
        for ep in range(E.EPI_COUNT):
            if ep in (E.EPI_BIAS_DROP_RES_LN, E.EPI_CE_STATS, E.EPI_CE_GRAD):
                continue
            for bn in (128, 256):
                f32 = ep in (E.EPI_ATOMIC_F32, E.EPI_STORE_F32)
                kw = dict(a_mn=bool(am), b_mn=bool(bm), epilogue=ep, block_n=bn,
                          out=torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16))
                if ep in E.EPI_HAS_BIAS:
                    kw["bias"] = torch.randn(N, device=dev)
                if ep in E.EPI_HAS_RES:
                    kw["res"] = torch.randn(M, N, device=dev).bfloat16()
                if ep == E.EPI_BIAS_DROP_RES:
                    kw["drop"] = _lib.dropout_t(seed, 3, 0.1)
                if ep == E.EPI_ATOMIC_F32:
                    kw["splits"] = 2
                if ep == E.EPI_MUL:
                    kw["colsum"] = torch.zeros(N, device=dev)
                ops.gemm(a, b, **kw)
    # LayerNorm epilogue: clusters of 1, 3 and 6 CTAs exchanging row statistics over DSMEM
    for N in (128, 384, 768):
        a = torch.randn(M, K, device=dev).bfloat16()
        b = torch.randn(N, K, device=dev).bfloat16()
        ops.gemm(a, b, epilogue=E.EPI_BIAS_DROP_RES_LN, bias=torch.randn(N, device=dev),
                 res=torch.randn(M, N, device=dev).bfloat16(), drop=_lib.dropout_t(seed, 3, 0.1),
                 ln=(torch.ones(N, device=dev), torch.zeros(N, device=dev), 1e-12,
                     torch.empty(M, device=dev), torch.empty(M, device=dev)))
    # vocabulary GEMM + cross entropy epilogues (ragged last column tile), forward and backward
    from meme_challenge_b200 import functional as F_
    xh = torch.randn(45, 64, device=dev).bfloat16().requires_grad_(True)
    Wv = torch.nn.Parameter(torch.randn(515, 64, device=dev) * 0.05)
    bv = torch.nn.Parameter(torch.zeros(515, device=dev))
    F_.vocab_cross_entropy(xh, Wv, bv, torch.randint(0, 515, (45,), device=dev)).sum().backward()
    # reduce step of the copy-engine gradient exchange
    import ctypes as C
    own = torch.randn(4096, device=dev).bfloat16()
    peers = [torch.randn(4096, device=dev).bfloat16() for _ in range(7)]
    ops._call("b200u_slice_sum_bf16", ops.P(own), (C.c_void_p * 7)(*[t.data_ptr() for t in peers]), 7, C.c_size_t(4096))
    torch.cuda.synchronize()
    print("gemm ok", flush=True)

if want("attention"):
    for (B, L, heads) in ((2, 164, 2), (2, 200, 2), (3, 37, 1)):
        H = heads * 64
        M = B * L
        qkv = (torch.randn(M, 3 * H, device=dev) * 0.5).bfloat16()
        mask = torch.zeros(B, L, device=dev)
        mask[:, L - 5:] = -10000.0
        d = _lib.dropout_t(seed, 5, 0.1)
        ctx, lse = ops.attention_fwd(qkv, mask, B, L, heads, H, drop=d)
        dctx = torch.randn(M, H, device=dev).bfloat16()
        dbias = torch.zeros(3 * H, device=dev)
        ops.attention_bwd(qkv, mask, ctx, dctx, lse, B, L, heads, H, drop=d, dbias_qkv=dbias)
    torch.cuda.synchronize()
    print("attention ok", flush=True)

if want("layernorm"):
    M, H = 333, 768
    x = torch.randn(M, H, device=dev).bfloat16()
    dy = torch.randn(M, H, device=dev).bfloat16()
    gamma = torch.ones(H, device=dev)
    beta = torch.zeros(H, device=dev)
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-12)
    dg, db_, dbias = (torch.zeros(H, device=dev) for _ in range(3))
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg, db_, dz=True, dbias=dbias, drop=_lib.dropout_t(seed, 3, 0.1))
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg, db_)
    big = torch.randn(M, 3072, device=dev).bfloat16()
    ops.colsum_accum(big, torch.zeros(3072, device=dev))
    torch.cuda.synchronize()
    print("layernorm ok", flush=True)

if want("model"):
    # tiny MemeUniter fwd+bwd + one fused optimizer step: embeddings, gather, heads, optimizer kernels
    from meme_challenge_b200.data.synthetic import synth_batch
    from meme_challenge_b200.model.meme_uniter import MemeUniter
    from meme_challenge_b200.model.model import UniterConfig, UniterModel
    from meme_challenge_b200.train import TrainStep
    cfg = UniterConfig.from_dict(dict(vocab_size=512, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                      intermediate_size=256, hidden_act="gelu", hidden_dropout_prob=0.1,
                                      attention_probs_dropout_prob=0.1, max_position_embeddings=64,
                                      type_vocab_size=2, initializer_range=0.02))
    m = MemeUniter(UniterModel(cfg, 64), 128, 1).to(dev).train()
    ts = TrainStep(m, gradient_accumulation=2)
    bs = []
    for i in range(2):
        b = synth_batch(3, 12, 10, seed=11 + i, variable=True, img_dim=64, vocab=512, min_txt=3, min_bb=2)
        b = {k: v.to(dev) for k, v in b.items() if torch.is_tensor(v)}
        b["labels"] = b["labels"].float()
        bs.append(b)
    ts.pipeline = False
    ts.step(bs)
    torch.cuda.synchronize()
    print("model ok", flush=True)

if want("ot"):
    from meme_challenge_b200.model import ot
    x = torch.randn(2, 12, 64, device=dev)
    y = torch.randn(2, 10, 64, device=dev)
    xp = torch.zeros(2, 12, dtype=torch.bool, device=dev)
    yp = torch.zeros(2, 10, dtype=torch.bool, device=dev)
    xp[:, 9:] = True
    ot.optimal_transport_dist(x, y, xp, yp)
    torch.cuda.synchronize()
    print("ot ok", flush=True)
print("done")
