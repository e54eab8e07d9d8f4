"""GPU (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle on the same seeded
inputs. Tolerances are BASELINE.json's: gather/mask bit-exact; logits max-abs <= 1e-2; loss rel
<= 1e-3 (bf16 compute / fp32 accumulate)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import uniter_oracle as O
from oracle.make_golden import IMG_DIM, TINY

DEV = "cuda"
LOGIT_TOL = 1e-2     # BASELINE.json north_star
LOSS_RTOL = 1e-3     # BASELINE.json north_star


def _require_gpu():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from meme_challenge_b200 import _lib
    import ctypes
    sm, ma, mi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.lib().b200u_device_info(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi)))
    assert ma.value == 10, "built for sm_100a, found cc %d.%d" % (ma.value, mi.value)


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


# ----------------------------------------------------------------------------------------------
# kernels
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("shape", [(128, 128, 64), (2624, 768, 768), (200, 136, 72), (768, 3072, 2624)])
def test_gemm_layouts(shape, a_mn, b_mn):
    _require_gpu()
    from meme_challenge_b200 import ops
    M, N, K = shape
    torch.manual_seed(1)
    a = torch.randn((K, M) if a_mn else (M, K), device=DEV).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device=DEV).bfloat16()
    ref = (a.float().t() if a_mn else a.float()).cpu() @ (b.float() if b_mn else b.float().t()).cpu()
    for bn in (128, 256):
        out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, block_n=bn).float().cpu()
        assert (out - ref).abs().max() <= 1e-2 * ref.abs().max()


def test_gemm_epilogues_and_splitk():
    _require_gpu()
    from meme_challenge_b200 import _lib, ops
    M, N, K = 520, 384, 320
    torch.manual_seed(2)
    a = torch.randn(M, K, device=DEV).bfloat16()
    b = torch.randn(N, K, device=DEV).bfloat16()
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV).bfloat16()
    ref = a.float().cpu() @ b.float().cpu().t()
    tol = 1e-2 * ref.abs().max()
    u, g = ops.gemm(a, b, bias=bias, epilogue=_lib.EPI_BIAS_GELU)
    assert (u.float().cpu() - (ref + bias.cpu())).abs().max() <= tol
    assert (g.float().cpu() - O.gelu(u.float().cpu())).abs().max() <= 1e-2 * O.gelu(ref).abs().max()
    out = ops.gemm(a, b, bias=bias, res=res, epilogue=_lib.EPI_BIAS_DROP_RES)
    assert (out.float().cpu() - (ref + bias.cpu() + res.float().cpu())).abs().max() <= tol
    x = res.float().cpu().requires_grad_(True)
    O.gelu(x).sum().backward()
    out = ops.gemm(a, b, res=res, epilogue=_lib.EPI_DGELU)
    assert (out.float().cpu() - ref * x.grad).abs().max() <= tol
    acc = torch.full((M, N), 2.0, device=DEV)
    ops.gemm(a, b, epilogue=_lib.EPI_ATOMIC_F32, out=acc, splits=3)
    assert (acc.cpu() - (ref + 2.0)).abs().max() <= 2e-3 * ref.abs().max()
    # tcgen05 path and the SIMT debug path share the epilogue: results agree to fp32 rounding
    o0 = ops.gemm(a, b, bias=bias, epilogue=_lib.EPI_STORE_F32, impl=0)
    o1 = ops.gemm(a, b, bias=bias, epilogue=_lib.EPI_STORE_F32, impl=1)
    assert (o0 - o1).abs().max() <= 1e-3 * ref.abs().max()


def test_gemm_gelu_derivative_and_mul_colsum_epilogues():
    """EPI_BIAS_GELU_DG saves gelu'(u) next to gelu(u) (model/layer.py:139-142 forward) and EPI_MUL consumes it
    in the backward with the bias-gradient column sums folded into the same epilogue."""
    _require_gpu()
    from meme_challenge_b200 import _lib, ops
    M, N, K = 520, 384, 320
    torch.manual_seed(12)
    a = torch.randn(M, K, device=DEV).bfloat16()
    b = (torch.randn(N, K, device=DEV) * 0.1).bfloat16()
    bias = torch.randn(N, device=DEV) * 0.5
    u_ref = (a.float().cpu() @ b.float().cpu().t() + bias.cpu()).bfloat16().float()   # the kernel rounds u to bf16 first
    x = u_ref.clone().requires_grad_(True)
    gval = O.gelu(x)
    gval.sum().backward()
    for impl in (0, 1):
        for bn in (128, 256):
            dg, gv = ops.gemm(a, b, bias=bias, epilogue=_lib.EPI_BIAS_GELU_DG, block_n=bn, impl=impl)
            # u itself may differ by one bf16 ulp between accumulation orders: compare through smooth functions
            assert (gv.float().cpu() - gval.detach()).abs().max() <= 2e-2 * gval.abs().max()
            assert (dg.float().cpu() - x.grad).abs().max() <= 2e-2
    # backward epilogue: out = acc * R, colsum += column sums of the bf16 output
    r = torch.rand(M, N, device=DEV).bfloat16()
    ref = (a.float().cpu() @ b.float().cpu().t()) * r.float().cpu()
    for impl in (0, 1):
        for bn in (128, 256):
            cs = torch.full((N,), 3.0, device=DEV)
            out = ops.gemm(a, b, res=r, epilogue=_lib.EPI_MUL, colsum=cs, block_n=bn, impl=impl)
            assert (out.float().cpu() - ref).abs().max() <= 1e-2 * ref.abs().max()
            want = out.float().sum(0).cpu() + 3.0
            assert (cs.cpu() - want).abs().max() <= 1e-3 * want.abs().max() + 1e-3
    # B stored [K, N] (the dgrad layout the BertLayer backward uses)
    bt = b.t().contiguous()
    cs = torch.zeros(N, device=DEV)
    out = ops.gemm(a, bt, b_mn=True, res=r, epilogue=_lib.EPI_MUL, colsum=cs)
    assert (out.float().cpu() - ref).abs().max() <= 1e-2 * ref.abs().max()
    assert (cs.cpu() - out.float().sum(0).cpu()).abs().max() <= 1e-3 * out.float().sum(0).abs().max().cpu() + 1e-3


@pytest.mark.parametrize("N,block_n", [(128, 0), (768, 0), (1024, 0), (256, 256), (512, 256), (768, 256)])
@pytest.mark.parametrize("M", [100, 2624, 5248])
def test_gemm_fused_layernorm_epilogue(M, N, block_n):
    """EPI_BIAS_DROP_RES_LN: y = acc + bias + R (bf16) and LayerNorm(y) over the full row from ONE launch, the
    row statistics exchanged between the N/128 CTAs of a cluster (model/layer.py:111-115,152-156)."""
    _require_gpu()
    from meme_challenge_b200 import _lib, ops
    K = 320
    torch.manual_seed(13)
    a = torch.randn(M, K, device=DEV).bfloat16()
    b = (torch.randn(N, K, device=DEV) * 0.1).bfloat16()
    bias = torch.randn(N, device=DEV) * 0.5
    res = (torch.randn(M, N, device=DEV) * 2 + 0.3).bfloat16()
    gamma = torch.randn(N, device=DEV) * 0.3 + 1
    beta = torch.randn(N, device=DEV) * 0.1
    mean = torch.empty(M, device=DEV)
    rstd = torch.empty(M, device=DEV)
    y, x = ops.gemm(a, b, bias=bias, res=res, epilogue=_lib.EPI_BIAS_DROP_RES_LN, block_n=block_n,
                    ln=(gamma, beta, 1e-12, mean, rstd))
    y_ref = a.float().cpu() @ b.float().cpu().t() + bias.cpu() + res.float().cpu()
    assert (y.float().cpu() - y_ref).abs().max() <= 1e-2 * y_ref.abs().max()
    # LayerNorm of the bf16 values the kernel stored (what the standalone kernel and the backward see)
    yb = y.float().cpu()
    x_ref = torch.nn.functional.layer_norm(yb, (N,), gamma.cpu(), beta.cpu(), 1e-12)
    # (bf16 output: one ulp is 2^-8 relative, and fp32 operation order may flip a rounding)
    assert ((x.float().cpu() - x_ref).abs() <= 1e-2 + 8e-3 * x_ref.abs()).all()
    assert (mean.cpu() - yb.mean(1)).abs().max() <= 1e-4
    rs = 1.0 / torch.sqrt(yb.var(1, unbiased=False) + 1e-12)
    assert ((rstd.cpu() - rs).abs() / rs).max() <= 1e-4
    # identical to the separate launches (GEMM + standalone LayerNorm) up to fp32 summation order
    y2 = ops.gemm(a, b, bias=bias, res=res, epilogue=_lib.EPI_BIAS_DROP_RES)
    assert torch.equal(y2, y)
    x2, mean2, rstd2 = ops.layernorm_fwd(y2, gamma, beta, 1e-12)
    assert ((x2.float() - x.float()).abs() <= 1e-2 + 8e-3 * x.float().abs()).all().item()
    assert (mean2 - mean).abs().max().item() <= 1e-5 and ((rstd2 - rstd).abs() / rstd2).max().item() <= 1e-5
    # dropout inside the fused epilogue draws the same mask as the unfused one (same counter stream)
    seed = torch.tensor([5], device=DEV, dtype=torch.int64)
    d = _lib.dropout_t(seed, 9, 0.1)
    yd, _ = ops.gemm(a, b, bias=bias, res=res, epilogue=_lib.EPI_BIAS_DROP_RES_LN, drop=d, block_n=block_n,
                     ln=(gamma, beta, 1e-12, None, None))
    yd2 = ops.gemm(a, b, bias=bias, res=res, epilogue=_lib.EPI_BIAS_DROP_RES, drop=d)
    assert torch.equal(yd, yd2)


def test_gemm_dropout_statistics():
    _require_gpu()
    from meme_challenge_b200 import _lib, ops
    M, N, K = 1024, 768, 64
    a = torch.ones(M, K, device=DEV).bfloat16()
    b = torch.ones(N, K, device=DEV).bfloat16()
    zero = torch.zeros(M, N, device=DEV).bfloat16()
    seed = torch.tensor([42], device=DEV, dtype=torch.int64)
    d = _lib.dropout_t(seed, 5, 0.1)
    o = ops.gemm(a, b, bias=torch.zeros(N, device=DEV), res=zero, epilogue=_lib.EPI_BIAS_DROP_RES, drop=d).float()
    keep = (o != 0).float().mean().item()
    assert abs(keep - 0.9) < 3e-3
    assert torch.allclose(o[o != 0], torch.tensor(64.0 / 0.9, device=DEV), rtol=1e-2)
    seed2 = torch.tensor([43], device=DEV, dtype=torch.int64)
    o2 = ops.gemm(a, b, bias=torch.zeros(N, device=DEV), res=zero, epilogue=_lib.EPI_BIAS_DROP_RES,
                  drop=_lib.dropout_t(seed2, 5, 0.1)).float()
    agree = ((o != 0) == (o2 != 0)).float().mean().item()
    assert abs(agree - 0.82) < 0.01  # independent masks: 0.9^2 + 0.1^2


@pytest.mark.parametrize("H", [128, 768, 1024])
def test_layernorm_fwd_bwd(H):
    _require_gpu()
    from meme_challenge_b200 import ops
    torch.manual_seed(3)
    M = 777
    x = (torch.randn(M, H) * 2 + 0.5)
    w, b = torch.randn(H) * 0.3 + 1, torch.randn(H) * 0.1
    dy = torch.randn(M, H)
    xb, dyb = x.bfloat16(), dy.bfloat16()
    xr = xb.float().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = O.layer_norm(xr, wr, br)
    yr.backward(dyb.float())
    y, mean, rstd = ops.layernorm_fwd(xb.to(DEV), w.to(DEV), b.to(DEV), 1e-12)
    assert (y.float().cpu() - yr.detach()).abs().max() < 3e-2
    dg, db_, dbias = (torch.zeros(H, device=DEV) for _ in range(3))
    dx, _ = ops.layernorm_bwd(dyb.to(DEV), xb.to(DEV), mean, rstd, w.to(DEV), dg, db_, dbias=dbias)
    assert _cos(dx.float().cpu(), xr.grad) > 0.9999
    assert (dx.float().cpu() - xr.grad).abs().max() < 2e-2 * xr.grad.abs().max()
    assert torch.allclose(dg.cpu(), wr.grad, rtol=1e-3, atol=1e-2)
    assert torch.allclose(db_.cpu(), br.grad, rtol=1e-3, atol=1e-2)
    assert torch.allclose(dbias.cpu(), dx.float().sum(0).cpu(), rtol=1e-3, atol=1e-2)


def test_gather_bit_exact_with_padded_positions():
    """K1: torch.equal with model/model.py:330-333 on identical inputs, identity tail included."""
    _require_gpu()
    from meme_challenge_b200 import ops
    torch.manual_seed(4)
    B, T, R, H = 5, 16, 20, 128
    tl = [16, 3, 9, 1, 12]
    nb = [20, 20, 4, 7, 13]
    am = O.get_attention_mask(tl, nb)
    L = am.shape[1]
    gi = O.get_gather_index(tl, nb, B, T, L)
    txt = torch.randn(B, T, H).bfloat16()
    img = torch.randn(B, R, H).bfloat16()
    want = torch.gather(torch.cat([txt, img], 1), 1, gi.unsqueeze(-1).expand(-1, -1, H))
    got = ops.gather_rows(txt.to(DEV), img.to(DEV), gi.to(DEV)).cpu()
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    # backward == scatter-add (duplicates in the identity tail add up)
    dout = torch.randn(B, L, H).bfloat16()
    cat = torch.cat([txt, img], 1).float().requires_grad_(True)
    torch.gather(cat, 1, gi.unsqueeze(-1).expand(-1, -1, H)).backward(dout.float())
    dtxt, dimg = ops.gather_rows_bwd(dout.to(DEV), gi.to(DEV), T, R)
    got_g = torch.cat([dtxt, dimg], 1).float().cpu()
    assert (got_g - cat.grad).abs().max() <= 2e-2 * cat.grad.abs().max()


def test_index_mask_on_device_bit_exact():
    _require_gpu()
    from meme_challenge_b200.utils.utils import get_attention_mask, get_gather_index
    tl, nb, T = [8, 64, 33, 1], [100, 36, 77, 50], 64
    am = get_attention_mask(tl, nb, device=DEV)
    gi = get_gather_index(tl, nb, 4, T, am.shape[1], device=DEV)
    assert torch.equal(am.cpu(), O.get_attention_mask(tl, nb))
    assert torch.equal(gi.cpu(), O.get_gather_index(tl, nb, 4, T, am.shape[1]))


@pytest.mark.parametrize("L,heads", [(164, 12), (76, 2), (40, 3), (200, 2)])
def test_attention_fwd_bwd(L, heads):
    _require_gpu()
    from meme_challenge_b200 import ops
    torch.manual_seed(5)
    B, H = 3, heads * 64
    qkv = (torch.randn(B * L, 3 * H) * 0.8).bfloat16()
    valid = torch.tensor([L, max(1, L // 2), max(1, L - 7)])
    mask01 = (torch.arange(L).unsqueeze(0) < valid.unsqueeze(1)).float()
    mask_add = (1.0 - mask01) * -10000.0
    dctx = torch.randn(B * L, H).bfloat16()

    x = qkv.float().requires_grad_(True)
    q, k, v = x.view(B, L, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0 + mask_add[:, None, None, :]
    ctx_ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, H)
    ctx_ref.backward(dctx.float())

    ctx, lse = ops.attention_fwd(qkv.to(DEV), mask_add.to(DEV), B, L, heads, H)
    assert (ctx.float().cpu() - ctx_ref.detach()).abs().max() < 2e-2
    dbias = torch.full((3 * H,), 0.5, device=DEV)   # accumulated into (+=), not overwritten
    dqkv = ops.attention_bwd(qkv.to(DEV), mask_add.to(DEV), ctx, dctx.to(DEV), lse, B, L, heads, H,
                             dbias_qkv=dbias)
    g = dqkv.float().cpu()
    # fused QKV bias gradient = column sums of the bf16 dqkv the same launch wrote
    want_db = dqkv.float().sum(0).cpu() + 0.5
    assert (dbias.cpu() - want_db).abs().max() <= 1e-3 * max(1.0, want_db.abs().max().item())
    assert _cos(g, x.grad) > 0.999
    assert (g - x.grad).abs().max() < 3e-2 * x.grad.abs().max()
    # masked keys receive exactly zero dK / dV, like the reference (exp underflow)
    gk = g.view(B, L, 3, H)[1, int(valid[1]):, 1:]
    assert gk.abs().max() == 0


def test_attention_dropout_is_consistent_between_fwd_and_bwd():
    """With dropout on, d(ctx)/d(V) must use the same mask the forward drew: check the linearity
    ctx(V1 + V2) == ctx(V1) + ctx(V2) for a fixed seed, and the backward against a finite
    difference through V."""
    _require_gpu()
    from meme_challenge_b200 import _lib, ops
    torch.manual_seed(6)
    B, L, heads = 2, 100, 2
    H = heads * 64
    qkv = (torch.randn(B * L, 3 * H) * 0.5).bfloat16().to(DEV)
    mask_add = torch.zeros(B, L, device=DEV)
    seed = torch.tensor([99], device=DEV, dtype=torch.int64)
    d = _lib.dropout_t(seed, 3, 0.1)
    ctx1, lse = ops.attention_fwd(qkv, mask_add, B, L, heads, H, drop=d)
    ctx2, _ = ops.attention_fwd(qkv, mask_add, B, L, heads, H, drop=d)
    assert torch.equal(ctx1, ctx2)  # same seed -> same mask
    ctx0, _ = ops.attention_fwd(qkv, mask_add, B, L, heads, H)
    assert not torch.equal(ctx0, ctx1)
    # E[dropout(P)] = P: averaged over rows the dropped context stays close to the clean one
    assert (ctx1.float().mean(0) - ctx0.float().mean(0)).abs().max() < 0.05
    # backward: dV = Pdᵀ·dO. With dO = one-hot rows, compare against forward differences in V.
    dctx = torch.zeros(B * L, H, device=DEV).bfloat16()
    dctx[5, 7] = 1.0
    dqkv = ops.attention_bwd(qkv, mask_add, ctx1, dctx, lse, B, L, heads, H, drop=d).float()
    dv = dqkv[:L, 2 * H + 7]          # d ctx[5, 7] / d V[j, 7] for sample 0, head 0  == Pd[5, j]
    q2 = qkv.clone().float()
    q2[:L, 2 * H:2 * H + 64] = 0
    q2[:L, 2 * H + 7] = torch.arange(L, device=DEV).float() % 2  # V[:, 7] = parity pattern
    c, _ = ops.attention_fwd(q2.bfloat16(), mask_add, B, L, heads, H, drop=d)
    want = (dv * (torch.arange(L, device=DEV).float() % 2)).sum()
    assert abs(c[5, 7].float().item() - want.item()) < 2e-2


def test_out_of_range_indices_are_flagged_not_dereferenced():
    """nn.Embedding / torch.gather raise on an index outside the table (model/model.py:232-245, 329-333). The
    kernels never dereference such an index: they flag it in the device's input-error word and substitute a safe
    row; check_input_errors() then raises. int32 ids (which nn.Embedding accepts) are converted, not misread."""
    _require_gpu()
    from meme_challenge_b200 import functional as F_, ops
    from meme_challenge_b200._lib import B200UError
    F_.check_input_errors()          # start clean
    b = O.synth_batch(2, 12, 10, seed=5, img_dim=IMG_DIM, vocab=TINY["vocab_size"], min_txt=2, min_bb=2)
    m = _build(TINY, IMG_DIM).eval()
    ref = m(**_kw(b)).detach().clone()
    F_.check_input_errors()
    # int32 ids give the same logits as int64
    kw = _kw(b)
    kw["input_ids"] = kw["input_ids"].int()
    kw["position_ids"] = kw["position_ids"].int()
    kw["gather_index"] = kw["gather_index"].int()
    assert torch.equal(m(**kw).detach(), ref)
    F_.check_input_errors()
    # a word id beyond the vocabulary, forward AND backward: flagged, no fault, other gradients untouched
    kw = _kw(b)
    kw["input_ids"] = kw["input_ids"].clone()
    kw["input_ids"][0, 1] = TINY["vocab_size"] + 7
    m.train()
    m(**kw).sum().backward()
    with pytest.raises(B200UError, match="word-embedding"):
        F_.check_input_errors()
    F_.check_input_errors()          # the check cleared the word
    # position id / gather index out of range
    kw = _kw(b)
    kw["position_ids"] = kw["position_ids"].clone()
    kw["position_ids"][0, 0] = -3
    m.eval()
    m(**kw)
    with pytest.raises(B200UError, match="position"):
        F_.check_input_errors()
    kw = _kw(b)
    kw["gather_index"] = kw["gather_index"].clone()
    kw["gather_index"][1, 2] = 10 ** 6
    m(**kw)
    with pytest.raises(B200UError, match="gather_index"):
        F_.check_input_errors()
    # float ids are rejected on the host
    kw = _kw(b)
    kw["input_ids"] = kw["input_ids"].float()
    with pytest.raises(B200UError, match="integer"):
        m(**kw)
    # scatter backward with a bad id: the table's neighbours in the flat gradient stay untouched
    d = torch.ones(4, 64, device=DEV).bfloat16()
    guard = torch.zeros(3, 8, 64, device=DEV)
    ids = torch.tensor([[0, 9, -1, 7]], device=DEV).t().contiguous()   # B = 4, T = 1; ids 9 and -1 are out of range
    import ctypes as C
    ops._call("b200u_embedding_scatter_add", ops.P(d), ops.P(ids), 1, 1, C.c_longlong(0), ops.P(guard[1]), 4, 64,
              C.c_longlong(-100), C.c_longlong(8))
    with pytest.raises(B200UError, match="gradient row"):
        F_.check_input_errors()
    assert guard[0].abs().sum().item() == 0 and guard[2].abs().sum().item() == 0
    assert guard[1].sum().item() == 2 * 64      # rows 0 and 7 only


@pytest.mark.parametrize("nsrc,n", [(1, 4096), (7, 8 * 1000 + 8), (15, 64)])
def test_slice_sum_bf16_matches_fp32_sum_in_order(nsrc, n):
    """b200u_slice_sum_bf16 (reduce step of the copy-engine gradient exchange): fp32 accumulation of the own slice
    and the peers' copies in the order given, one rounding to bf16 -> bit-exact against the same sum in torch."""
    _require_gpu()
    import ctypes as C
    from meme_challenge_b200 import ops
    torch.manual_seed(nsrc)
    own = torch.randn(n, device=DEV).bfloat16()
    srcs = [torch.randn(n, device=DEV).bfloat16() for _ in range(nsrc)]
    want = own.float()
    for t in srcs:
        want = want + t.float()
    want = want.bfloat16()
    arr = (C.c_void_p * nsrc)(*[t.data_ptr() for t in srcs])
    ops._call("b200u_slice_sum_bf16", ops.P(own), arr, nsrc, C.c_size_t(n))
    assert torch.equal(own.view(torch.int16), want.view(torch.int16))


# ----------------------------------------------------------------------------------------------
# whole model against the reference golden vectors and the oracle
# ----------------------------------------------------------------------------------------------
REL_L2_BOUND = 0.04       # matrices (weights, embedding tables); measured worst 0.029 (ragged C2 batch), median 0.012-0.020
REL_L2_BOUND_VEC = 0.05   # bias and LayerNorm vectors; measured worst 0.026


def _build(cfg_dict, img_dim, sd=None, seed=0):
    from meme_challenge_b200.model.meme_uniter import MemeUniter
    from meme_challenge_b200.model.model import UniterConfig, UniterModel
    torch.manual_seed(seed)
    cfg = UniterConfig.from_dict(cfg_dict)
    m = MemeUniter(UniterModel(cfg, img_dim), cfg.hidden_size, 1)
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    return m.to(DEV)


def _kw(b, dev=DEV):
    return dict(input_ids=b["input_ids"].to(dev), position_ids=b["position_ids"].to(dev),
                img_feat=b["img_feat"].to(dev), img_pos_feat=b["img_pos_feat"].to(dev),
                attention_mask=b["attn_mask"].to(dev), gather_index=b["gather_index"].to(dev),
                output_all_encoded_layers=False)


def test_tiny_model_against_reference_golden(golden_dir):
    """Committed fixtures from the unmodified reference: logits, loss and every parameter grad."""
    _require_gpu()
    g = np.load(os.path.join(golden_dir, "tiny_meme_uniter.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    b = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in.")}
    m = _build(TINY, IMG_DIM, sd).eval()
    logits = m(**_kw(b))
    assert logits.dtype == torch.float32 and tuple(logits.shape) == (4, 1)
    assert (logits.detach().cpu().numpy() - g["logits"]).__abs__().max() <= LOGIT_TOL
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([1.8], device=DEV))(
        logits.squeeze(1), b["labels"].float().to(DEV))
    assert abs(loss.item() - float(g["loss"])) <= LOSS_RTOL * abs(float(g["loss"]))
    loss.backward()
    worst = 1.0
    for n, p in m.named_parameters():
        key = "grad." + n
        if key not in g.files:
            assert p.grad is None or p.grad.abs().max() == 0, n
            continue
        want = torch.from_numpy(g[key])
        got = p.grad.detach().cpu()
        if want.abs().max() < 1e-7:
            continue
        c = _cos(got, want)
        worst = min(worst, c)
        assert c > 0.99, (n, c)
    assert worst > 0.99
    # intermediate tensors: embedding output and per-layer outputs on valid rows
    layers = m.uniter_model(**{**_kw(b), "output_all_encoded_layers": True})
    assert len(layers) == 2
    valid = b["attn_mask"].bool()
    for got, key in ((layers[0], "layer0_out"), (layers[1], "layer1_out")):
        d = (got.float().cpu() - torch.from_numpy(g[key])).abs()
        assert d[valid].max() < 5e-2


def _oracle_run(m, cfg_dict, b, pos_wt=1.8, want_grads=True, device="cpu"):
    """The fp32 oracle on the model's weights. device="cuda": the same restatement evaluated by torch's fp32 CUDA
    kernels (TF32 off) so full-depth large configurations stay in seconds; results come back on the CPU."""
    sd = {k: v.detach().float().to(device).clone().requires_grad_(want_grads) for k, v in m.state_dict().items()}
    kw = dict(input_ids=b["input_ids"], position_ids=b["position_ids"], img_feat=b["img_feat"],
              img_pos_feat=b["img_pos_feat"], attention_mask=b["attn_mask"], gather_index=b["gather_index"])
    kw = {k: v.to(device) for k, v in kw.items()}
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.set_grad_enabled(want_grads):
            logits = O.meme_uniter_forward(sd, cfg_dict, **kw)
            loss = O.bce_loss(logits, b["labels"].to(device), pos_wt)
        if want_grads:
            loss.backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    if device != "cpu":
        class _G(object):
            def __init__(self, g):
                self.grad = g
        sd = {k: _G(None if v.grad is None else v.grad.cpu()) for k, v in sd.items()}
    return logits.detach().cpu(), loss.detach().cpu(), sd


BASE = dict(vocab_size=28996, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
            intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
            attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
            initializer_range=0.02)


def test_base_c1_inference_logits():
    """BASELINE config 1: UNITER-base inference, B=16, 36 regions, 40 tokens (L=76)."""
    _require_gpu()
    b = O.synth_batch(16, 40, 36, seed=1234)
    m = _build(BASE, 2048).eval()
    with torch.no_grad():
        logits = m(**_kw(b)).cpu()
    ref, _, _ = _oracle_run(m, BASE, b, want_grads=False)
    assert (logits - ref).abs().max() <= LOGIT_TOL


@pytest.mark.parametrize("variable", [False, True])
def test_base_c2_fwd_bwd_against_oracle(variable):
    """BASELINE config 2 shape: B=16, 100 regions, 64 tokens (L=164), fwd+bwd, pos_wt 1.8 BCE
    (dropout off for parity). Fixed and ragged (36-100 regions) batches."""
    _require_gpu()
    b = O.synth_batch(16, 64, 100, seed=1234 + int(variable), variable=variable)
    m = _build(BASE, 2048).eval()
    logits = m(**_kw(b))
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([1.8], device=DEV))(
        logits.squeeze(1), b["labels"].float().to(DEV))
    loss.backward()
    ref_logits, ref_loss, sd = _oracle_run(m, BASE, b)
    assert (logits.detach().cpu() - ref_logits).abs().max() <= LOGIT_TOL
    assert abs(loss.item() - ref_loss.item()) <= LOSS_RTOL * abs(ref_loss.item())
    # per-tensor relative L2 error of every gradient against the fp32 oracle. The path computes in bf16 (operands
    # rounded to 2^-9 relative, fp32 accumulation), so a tensor's error is a few bf16 ulps amplified by the depth of
    # the chain behind it: bounds are REL_L2_BOUND for matrices and the looser REL_L2_BOUND_VEC for bias / LayerNorm
    # vectors, whose few hundred entries are sums of strongly cancelling terms.
    bad, worst = [], []
    for n, p in m.named_parameters():
        want = sd[n].grad
        if want is None or want.abs().max() < 1e-8:
            continue
        got_g = p.grad.detach().cpu().float()
        rel = ((got_g - want).norm() / want.norm()).item()
        worst.append((rel, n))
        bound = REL_L2_BOUND if want.dim() >= 2 else REL_L2_BOUND_VEC
        if rel > bound or _cos(got_g, want) < 0.98:
            bad.append((n, rel))
    worst.sort(reverse=True)
    if os.environ.get("B200U_PRINT_REL"):
        print("per-tensor rel-L2 (worst 12):", [(round(r, 4), n) for r, n in worst[:12]])
        print("median rel-L2:", worst[len(worst) // 2][0])
    assert not bad, (bad[:8], worst[:3])
    # whole-gradient agreement
    names = [n for n, p in m.named_parameters() if sd[n].grad is not None]
    got = torch.cat([dict(m.named_parameters())[n].grad.detach().cpu().flatten() for n in names])
    want = torch.cat([sd[n].grad.flatten() for n in names])
    assert _cos(got, want) > 0.995
    assert abs(got.norm().item() / want.norm().item() - 1) < 2e-2


def test_state_dict_roundtrip_with_oracle():
    """save -> oracle load -> same logits; and fused buffers still export separate q/k/v."""
    _require_gpu()
    m = _build(TINY, IMG_DIM).eval()
    b = O.synth_batch(3, 10, 6, seed=5, variable=True, img_dim=IMG_DIM, vocab=TINY["vocab_size"], min_txt=2, min_bb=2)
    with torch.no_grad():
        l0 = m(**_kw(b)).cpu()
    sd = {k: v.cpu().clone() for k, v in m.state_dict().items()}
    assert "uniter_model.encoder.layer.0.attention.self.key.weight" in sd
    ref, _, _ = _oracle_run(m, TINY, b, want_grads=False)
    assert (l0 - ref).abs().max() <= LOGIT_TOL
    m2 = _build(TINY, IMG_DIM, sd, seed=123).eval()
    with torch.no_grad():
        l1 = m2(**_kw(b)).cpu()
    assert torch.equal(l0, l1)


def test_training_mode_dropout_runs_and_differs():
    _require_gpu()
    m = _build(TINY, IMG_DIM).train()
    b = O.synth_batch(4, 12, 10, seed=2, img_dim=IMG_DIM, vocab=TINY["vocab_size"], min_txt=2, min_bb=2)
    l1 = m(**_kw(b))
    l2 = m(**_kw(b))
    assert not torch.equal(l1, l2)  # fresh masks per forward
    l2.sum().backward()
    g = m.uniter_model.encoder.layer[0].attention.self.query.weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().max() > 0


def test_second_backward_through_released_activations_raises():
    """The layer's saved activations are raw device buffers released after the first backward: a second backward
    (retain_graph=True) must fail loudly instead of reading freed memory."""
    _require_gpu()
    m = _build(TINY, IMG_DIM).eval()
    b = O.synth_batch(2, 12, 10, seed=3, img_dim=IMG_DIM, vocab=TINY["vocab_size"], min_txt=2, min_bb=2)
    out = m(**_kw(b)).sum()
    out.backward(retain_graph=True)
    with pytest.raises(Exception, match="already run"):
        out.backward()


def test_grad_accumulation_and_zero_grad_set_to_none():
    """Two backward passes accumulate like the reference's `elif grad_step: loss.backward()`
    (train_template.py:108-109); zero_grad(set_to_none=True) is survived."""
    _require_gpu()
    m = _build(TINY, IMG_DIM).eval()
    b = O.synth_batch(4, 12, 10, seed=3, img_dim=IMG_DIM, vocab=TINY["vocab_size"], min_txt=2, min_bb=2)
    m(**_kw(b)).sum().backward()
    w = m.uniter_model.encoder.layer[1].output.dense.weight
    g1 = w.grad.clone()
    m(**_kw(b)).sum().backward()
    assert torch.allclose(w.grad, 2 * g1, rtol=1e-3, atol=1e-6)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    opt.zero_grad(set_to_none=True)
    m(**_kw(b)).sum().backward()
    assert torch.allclose(w.grad, g1, rtol=1e-3, atol=1e-6)
    opt.step()
    with torch.no_grad():
        l_after = m(**_kw(b))
    assert torch.isfinite(l_after).all()


LARGE_SHAPES = dict(vocab_size=28996, hidden_size=1024, num_hidden_layers=2, num_attention_heads=16,
                    intermediate_size=4096, hidden_act="gelu", hidden_dropout_prob=0.1,
                    attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
                    initializer_range=0.02)


def test_large_c4_shapes_fwd_bwd_against_oracle():
    """BASELINE config 4 layer shapes (config/uniter-large.json: H=1024, I=4096, 16 heads of 64), two
    layers deep so the CPU oracle stays in seconds: ragged batch, fwd + bwd, same tolerances as C2."""
    _require_gpu()
    b = O.synth_batch(6, 24, 40, seed=77, variable=True, min_txt=4, min_bb=10)
    m = _build(LARGE_SHAPES, 2048).eval()
    logits = m(**_kw(b))
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([1.8], device=DEV))(
        logits.squeeze(1), b["labels"].float().to(DEV))
    loss.backward()
    ref_logits, ref_loss, sd = _oracle_run(m, LARGE_SHAPES, b)
    assert (logits.detach().cpu() - ref_logits).abs().max() <= LOGIT_TOL
    assert abs(loss.item() - ref_loss.item()) <= LOSS_RTOL * abs(ref_loss.item())
    names = [n for n, p in m.named_parameters() if sd[n].grad is not None and p.grad is not None]
    got = torch.cat([dict(m.named_parameters())[n].grad.detach().cpu().flatten() for n in names])
    want = torch.cat([sd[n].grad.flatten() for n in names])
    assert _cos(got, want) > 0.995
    for n in ("uniter_model.encoder.layer.1.intermediate.dense.weight",
              "uniter_model.encoder.layer.0.attention.self.value.weight",
              "uniter_model.encoder.layer.0.output.LayerNorm.weight",
              "uniter_model.img_embeddings.img_linear.weight"):
        assert _cos(dict(m.named_parameters())[n].grad.detach().cpu(), sd[n].grad) > 0.98, n


LARGE = dict(vocab_size=28996, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
             intermediate_size=4096, hidden_act="gelu", hidden_dropout_prob=0.1,
             attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
             initializer_range=0.02)


def test_large_c4_full_depth_fwd_bwd_against_oracle():
    """BASELINE config 4 at FULL depth (config/uniter-large.json: 24 layers, H=1024, I=4096, 16 heads) on the C2
    batch shape (16 memes, 64 tokens, 36-100 regions): fwd + bwd against the fp32 oracle evaluated on the GPU.
    Logit / loss tolerances as C2; per-tensor gradient rel-L2 bounds for twice the depth."""
    _require_gpu()
    b = O.synth_batch(16, 64, 100, seed=4321, variable=True)
    m = _build(LARGE, 2048).eval()
    logits = m(**_kw(b))
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([1.8], device=DEV))(
        logits.squeeze(1), b["labels"].float().to(DEV))
    loss.backward()
    ref_logits, ref_loss, sd = _oracle_run(m, LARGE, b, device=DEV)
    assert (logits.detach().cpu() - ref_logits).abs().max() <= LOGIT_TOL
    assert abs(loss.item() - ref_loss.item()) <= LOSS_RTOL * abs(ref_loss.item())
    worst = []
    for n, p in m.named_parameters():
        want = sd[n].grad
        if want is None or p.grad is None or want.abs().max() < 1e-8:
            continue
        got_g = p.grad.detach().cpu().float()
        worst.append((((got_g - want).norm() / want.norm()).item(), n))
    worst.sort(reverse=True)
    if os.environ.get("B200U_PRINT_REL"):
        print("large: per-tensor rel-L2 (worst 8):", [(round(r, 4), n) for r, n in worst[:8]], "median", worst[len(worst) // 2][0])
    # measured: worst 0.116 (query / key weights of the last layers, whose gradients pass through 24 softmaxes'
    # worth of bf16 rounding), median 0.03
    assert len(worst) > 370 and worst[0][0] <= 0.15 and worst[len(worst) // 2][0] <= 0.06, (worst[:5], worst[len(worst) // 2])
    names = [n for n, p in m.named_parameters() if sd[n].grad is not None and p.grad is not None]
    got = torch.cat([dict(m.named_parameters())[n].grad.detach().cpu().flatten() for n in names])
    want = torch.cat([sd[n].grad.flatten() for n in names])
    assert _cos(got, want) > 0.995
    assert abs(got.norm().item() / want.norm().item() - 1) < 2e-2


def test_train_step_matches_reference_optimizer_semantics():
    """SURVEY §8a row 16 (train_template.py:89-109, utils/optim_utils.py:16-46): gradients of the
    accumulation window are summed, divided by the window length, clipped to max_grad_norm (global L2),
    then torch.optim.Adam with L2 weight decay on everything whose name has no 'bias' / 'LayerNorm.bias' /
    'LayerNorm.weight', and zeroed. The fused optimizer (3 launches over the flat buffers) must land on
    the same parameters as torch.optim.Adam fed the SAME gradients, for two consecutive steps (bias
    correction, moment state), and must refresh the bf16 weight shadow the GEMMs read."""
    _require_gpu()
    from meme_challenge_b200.train import TrainStep, NO_DECAY
    cfg = dict(TINY)
    cfg["hidden_dropout_prob"] = 0.0
    cfg["attention_probs_dropout_prob"] = 0.0
    m = _build(cfg, IMG_DIM).train()
    lr, wd, accum, max_norm = 1e-3, 1e-3, 2, 0.01     # max_norm small enough that clipping is active
    ts = TrainStep(m, lr=lr, weight_decay=wd, gradient_accumulation=accum, max_grad_norm=max_norm, pos_wt=1.8)
    names = [n for n, _ in m.named_parameters()]
    ref_p = [p.detach().clone().requires_grad_(True) for _, p in m.named_parameters()]
    decay = [p for n, p in zip(names, ref_p) if not any(nd in n for nd in NO_DECAY)]
    no_decay = [p for n, p in zip(names, ref_p) if any(nd in n for nd in NO_DECAY)]
    assert decay and no_decay
    opt = torch.optim.Adam([{"params": decay, "weight_decay": wd}, {"params": no_decay, "weight_decay": 0.0}], lr=lr)
    for step in range(2):
        batches = []
        for i in range(accum):
            b = O.synth_batch(4, 12, 10, seed=40 + 2 * step + i, variable=True, img_dim=IMG_DIM,
                              vocab=TINY["vocab_size"], min_txt=2, min_bb=2)
            d = {k: v.to(DEV) for k, v in b.items() if torch.is_tensor(v)}
            d["labels"] = b["labels"].float().to(DEV)
            batches.append(d)
        for i, d in enumerate(batches):
            ts.micro_step(d, last=(i == accum - 1))
        grads = [p.grad.detach().clone() for _, p in m.named_parameters()]   # summed over the window
        ts.optimizer_step()
        # reference on the same gradients
        for rp, g in zip(ref_p, grads):
            # parameters the step never touches have .grad None in the reference (loss.backward() leaves them
            # alone, e.g. img_embeddings.mask_embedding) and torch.optim.Adam skips them: no decay, no moments
            rp.grad = (g / accum) if g.abs().sum() > 0 else None
        total = torch.nn.utils.clip_grad_norm_([rp for rp in ref_p if rp.grad is not None], max_norm)
        assert total > max_norm, "pick max_norm so that the clip is exercised"
        opt.step()
        assert abs(ts.gnorm.item() - total.item()) <= 1e-4 * total.item()
        worst = 0.0
        for (n, p), rp in zip(m.named_parameters(), ref_p):
            worst = max(worst, (p.detach() - rp.detach()).abs().max().item())
            assert torch.allclose(p.detach(), rp.detach(), rtol=1e-5, atol=2e-7), (step, n)
            assert p.grad is None or p.grad.abs().max() == 0, (step, n)       # zero_grad
        # the bf16 shadow (what the next forward's GEMMs read) follows the fp32 master weights
        w = m.uniter_model.encoder.layer[0].intermediate.dense.weight
        assert torch.equal(ts.store.w16(w), w.detach().bfloat16())


def test_pipelined_window_equals_sequential_window():
    """TrainStep.step software-pipelines the micro-batches of an accumulation window over two streams
    (forward i+1 beside backward i) and runs the weight-gradient GEMMs on a side stream: the parameters
    after two optimizer steps must equal those of the plain sequential, single-stream order (fp32
    accumulation order is the only difference), eagerly and through CUDA-graph replay."""
    _require_gpu()
    from meme_challenge_b200 import _lib
    from meme_challenge_b200.train import TrainStep
    cfg = dict(TINY)
    cfg["hidden_dropout_prob"] = 0.0
    cfg["attention_probs_dropout_prob"] = 0.0

    def batches(step):
        out = []
        for i in range(2):
            b = O.synth_batch(4, 12, 10, seed=60 + 2 * step + i, img_dim=IMG_DIM, vocab=TINY["vocab_size"],
                              min_txt=2, min_bb=2)
            d = {k: v.to(DEV) for k, v in b.items() if torch.is_tensor(v)}
            d["labels"] = b["labels"].float().to(DEV)
            out.append(d)
        return out

    def run(pipeline, graph):
        _lib.lib().b200u_set_bwd_streams(1 if pipeline else 0)
        try:
            m = _build(cfg, IMG_DIM, seed=3).train()
            ts = TrainStep(m, lr=1e-3, weight_decay=1e-3, gradient_accumulation=2, max_grad_norm=5.0, pos_wt=1.8)
            ts.pipeline = pipeline
            losses = []
            if graph:
                ts.capture(batches(0), warmup=0)
                for step in range(2):
                    ts.load_static(batches(step))
                    outs = ts.replay()
                    losses.append([float(o[0].item()) for o in outs])
            else:
                for step in range(2):
                    outs = ts.step(batches(step))
                    losses.append([float(o[0].item()) for o in outs])
            torch.cuda.synchronize()
            return {n: p.detach().clone() for n, p in m.named_parameters()}, losses
        finally:
            _lib.lib().b200u_set_bwd_streams(1)

    ref_p, ref_l = run(False, False)
    for graph in (False, True):
        got_p, got_l = run(True, graph)
        assert np.allclose(np.array(got_l), np.array(ref_l), rtol=1e-5, atol=1e-6), (graph, got_l, ref_l)
        for n in ref_p:
            # Adam's first steps move every weight by ~lr regardless of gradient size, so a gradient
            # that differs in the last fp32 bits can move a near-zero-gradient weight differently:
            # compare with a tolerance of a small fraction of lr
            assert (got_p[n] - ref_p[n]).abs().max() <= 2e-4, (graph, n)
            assert _cos(got_p[n] - 0, ref_p[n] - 0) > 0.999999, (graph, n)


def test_fused_window_equals_sequential_window():
    """TrainStep(fuse_window=True) runs the micro-batches of an accumulation window as ONE pass over all their
    samples (ragged micro-batches padded to a common width with masked columns / identity gather tail): the
    per-micro-batch losses and the parameters after two optimizer steps must equal the sequential window's,
    eagerly and through CUDA-graph replay. Also the reference's first-step quirk (train_template.py:101-103):
    a window of ONE micro-batch still divides by the accumulation count."""
    _require_gpu()
    from meme_challenge_b200.train import TrainStep
    cfg = dict(TINY)
    cfg["hidden_dropout_prob"] = 0.0
    cfg["attention_probs_dropout_prob"] = 0.0

    def batches(step, variable):
        out = []
        for i in range(2):
            b = O.synth_batch(4, 12, 10, seed=80 + 2 * step + i, variable=variable, img_dim=IMG_DIM,
                              vocab=TINY["vocab_size"], min_txt=2, min_bb=2)
            d = {k: v.to(DEV) for k, v in b.items() if torch.is_tensor(v)}
            d["labels"] = b["labels"].float().to(DEV)
            out.append(d)
        return out

    def run(fuse, graph, variable):
        m = _build(cfg, IMG_DIM, seed=3).train()
        ts = TrainStep(m, lr=1e-3, weight_decay=1e-3, gradient_accumulation=2, max_grad_norm=5.0, pos_wt=1.8,
                       fuse_window=fuse)
        ts.pipeline = False
        losses = []
        if graph:
            ts.capture(batches(0, variable), warmup=0)
            for step in range(2):
                ts.load_static(batches(step, variable))
                outs = ts.replay()
                losses.append([float(o[0].item()) for o in outs])
        else:
            for step in range(2):
                outs = ts.step(batches(step, variable))
                losses.append([float(o[0].item()) for o in outs])
        torch.cuda.synchronize()
        return {n: p.detach().clone() for n, p in m.named_parameters()}, losses

    for variable, graph in ((True, False), (False, False), (False, True)):
        ref_p, ref_l = run(False, False, variable)
        got_p, got_l = run(True, graph, variable)
        assert np.allclose(np.array(got_l), np.array(ref_l), rtol=1e-4, atol=1e-5), (variable, graph, got_l, ref_l)
        for n in ref_p:
            assert (got_p[n] - ref_p[n]).abs().max() <= 2e-4, (variable, graph, n)
            assert _cos(got_p[n] - 0, ref_p[n] - 0) > 0.999999, (variable, graph, n)

    # first-step quirk: one micro-batch, gradient still divided by accum = 2
    m = _build(cfg, IMG_DIM, seed=3).train()
    ts = TrainStep(m, lr=1e-3, weight_decay=0.0, gradient_accumulation=2, max_grad_norm=1e9, pos_wt=1.8)
    b0 = batches(0, True)[:1]
    logits = m(**_kw(b0[0]))
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([1.8], device=DEV))(logits.squeeze(1), b0[0]["labels"])
    ts.store.zero_grad()
    loss.backward()
    want = ts.store.grad.norm().item() / 2.0     # every parameter gradient lives in the flat buffer
    ts.store.zero_grad()
    ts.step(b0)
    assert abs(ts.gnorm.item() - want) <= 1e-3 * want


def test_prefetcher_eval_path_and_optimizer_state_roundtrip(tmp_path):
    """data.pipeline.PinnedPrefetcher feeds device batches that equal the host batches; evaluate.predict runs the
    no-grad forward + sigmoid + BCE on them (probabilities equal the oracle's within the logit tolerance);
    TrainStep.state_dict() round-trips through load_state_dict() (moments, step, skip set) and a resumed run
    lands on the same parameters as an uninterrupted one."""
    _require_gpu()
    from meme_challenge_b200 import evaluate as E
    from meme_challenge_b200.data.pipeline import PinnedPrefetcher
    from meme_challenge_b200.train import TrainStep
    cfg = dict(TINY)
    cfg["hidden_dropout_prob"] = 0.0
    cfg["attention_probs_dropout_prob"] = 0.0
    host = []
    for i in range(5):
        b = O.synth_batch(4, 12, 10, seed=90 + i, variable=True, img_dim=IMG_DIM, vocab=TINY["vocab_size"], min_txt=2, min_bb=2)
        hb = {k: v for k, v in b.items() if torch.is_tensor(v)}
        hb["ids"] = torch.arange(4) + 10 * i
        host.append(hb)
    got = []
    for db in PinnedPrefetcher(iter(host), DEV, depth=2):
        got.append({k: v.clone() for k, v in db.items()})
    assert len(got) == 5
    for hb, db in zip(host, got):
        for k in hb:
            want = hb[k].float() if k == "labels" else hb[k]
            assert torch.equal(db[k].cpu(), want), k
    # evaluation path
    m = _build(cfg, IMG_DIM, seed=4).eval()
    probs, labels, loss, ids = E.predict(m, got, pos_wt=1.8)
    ref_p = []
    for hb in host:
        ref_logits, _, _ = _oracle_run(m, cfg, hb, want_grads=False)
        ref_p.append(torch.sigmoid(ref_logits.reshape(-1)))
    ref_p = torch.cat(ref_p)
    assert (probs.cpu() - ref_p).abs().max() <= 5e-3
    assert torch.equal(ids.cpu(), torch.cat([hb["ids"] for hb in host]))
    metrics = E.standard_metrics_binary(probs.cpu(), labels.cpu(), add_optimal_acc=True)
    assert 0.0 <= metrics["aucroc"] <= 1.0 and metrics["optimal_accuracy"] >= metrics["accuracy"] - 1e-9
    E.export_predictions(str(tmp_path / "preds.csv"), ids, probs, labels=labels)
    assert len(open(str(tmp_path / "preds.csv")).read().splitlines()) == 21
    # optimizer state: 2 steps + resume for 1 == 3 steps
    def window(step):
        return [got[(2 * step) % 5], got[(2 * step + 1) % 5]]
    ma = _build(cfg, IMG_DIM, seed=5).train()
    ta = TrainStep(ma, lr=1e-3, weight_decay=1e-3, gradient_accumulation=2)
    for st in range(3):
        ta.step(window(st))
    mb = _build(cfg, IMG_DIM, seed=5).train()
    tb = TrainStep(mb, lr=1e-3, weight_decay=1e-3, gradient_accumulation=2)
    for st in range(2):
        tb.step(window(st))
    sd_model = {k: v.clone() for k, v in mb.state_dict().items()}
    sd_opt = tb.state_dict()
    assert set(sd_opt["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd_opt["state"][0]["step"]) == 2.0
    mc = _build(cfg, IMG_DIM, sd_model, seed=99).train()
    tc = TrainStep(mc, lr=1e-3, weight_decay=1e-3, gradient_accumulation=2)
    tc.load_state_dict(sd_opt)
    tc.step(window(2))
    for (n, pa), (_, pc) in zip(ma.named_parameters(), mc.named_parameters()):
        assert torch.allclose(pa, pc, rtol=1e-5, atol=1e-7), n
    # parameters without a gradient (mask_embedding in fine-tuning) are left alone, like torch.optim.Adam
    assert "uniter_model.img_embeddings.mask_embedding.weight" in ta.skip
    w0 = _build(cfg, IMG_DIM, seed=5).uniter_model.img_embeddings.mask_embedding.weight
    assert torch.equal(ma.uniter_model.img_embeddings.mask_embedding.weight.detach().cpu(), w0.detach().cpu())


def test_pretrain_step_round_robin_decreases_losses():
    """PretrainStep (BASELINE config 5 driver): MLM / MRFR / ITM(+IPOT) round robin with the fused optimizer;
    a few steps on one repeated batch reduce every task's loss and keep all parameters finite."""
    _require_gpu()
    from meme_challenge_b200.data.synthetic import synth_pretrain_batch
    from meme_challenge_b200.model.model import UniterConfig
    from meme_challenge_b200.model.pretrain import UniterForPretraining
    from meme_challenge_b200.train import PretrainStep
    cfg = dict(TINY)
    cfg["hidden_dropout_prob"] = 0.0
    cfg["attention_probs_dropout_prob"] = 0.0
    torch.manual_seed(0)
    m = UniterForPretraining(UniterConfig.from_dict(cfg), IMG_DIM, 24).to(DEV).train()
    b = synth_pretrain_batch(4, 12, 10, seed=77, img_dim=IMG_DIM, vocab=TINY["vocab_size"], label_dim=24, min_txt=2, min_bb=2)
    d = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b.items()}
    d["ot_inputs"] = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b["ot_inputs"].items()}
    ts = PretrainStep(m, lr=2e-3, weight_decay=0.0, max_grad_norm=5.0)
    first, last = {}, {}
    for i in range(18):
        loss, task = ts.task_step(d)
        first.setdefault(task, float(loss.item()))
        last[task] = float(loss.item())
    assert set(first) == {"mlm", "mrfr", "itm"}
    for t in first:
        assert last[t] < first[t], (t, first[t], last[t])
    assert all(torch.isfinite(p).all() for p in m.parameters())
    assert m.last_ot_loss is not None   # the OT distances are computed (and dropped) like the reference does


def test_data_parallel_replicas_identical_and_equal_single_process():
    """one process per GPU (torchrun, NCCL), gradient buckets all-reduced from the backward
    hooks, sparse word-embedding row exchange. tools/dp_check.py asserts that all ranks end with bit-identical
    parameters and that they equal a single-process run over the union of the ranks' micro-batches — fp32 and
    bf16 gradient buckets, eager and CUDA-graph replay, pipelined and fused windows. Needs >= 2 GPUs."""
    _require_gpu()
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run under `gpurun --gpus 2`)")
    import subprocess
    import sys
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0 and "DP CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


# ----------------------------------------------------------------------------------------------
# optimal transport (model/ot.py): golden vectors from the unmodified reference + oracle
# ----------------------------------------------------------------------------------------------
def test_ot_against_reference_golden(golden_dir):
    """IPOT plan rel error <= 1e-3, OT distance rel <= 1e-3 (BASELINE.json north_star)."""
    _require_gpu()
    from meme_challenge_b200.model import ot
    g = np.load(os.path.join(golden_dir, "ot.npz"))
    txt, img = torch.from_numpy(g["txt"]).to(DEV), torch.from_numpy(g["img"]).to(DEV)
    txt_pad, img_pad = torch.from_numpy(g["txt_pad"]).to(DEV), torch.from_numpy(g["img_pad"]).to(DEV)
    cost = ot.cost_matrix_cosine(txt, img)
    assert (cost.cpu() - torch.from_numpy(g["cost"])).abs().max() < 1e-5
    dist, T, _ = ot.optimal_transport_plan(txt, img, txt_pad, img_pad)
    Tref = torch.from_numpy(g["T"])
    assert (T.cpu() - Tref).norm() / Tref.norm() <= 1e-3
    dref = torch.from_numpy(g["dist"])
    assert ((dist.cpu() - dref).abs() / dref.abs()).max() <= 1e-3
    # the reference-signature ipot() entry point
    joint = txt_pad.unsqueeze(-1) | img_pad.unsqueeze(-2)
    cm = cost.masked_fill(joint, 0)
    T2 = ot.ipot(cm, (6 - txt_pad.sum(1)).float(), txt_pad, (9 - img_pad.sum(1)).float(), img_pad, joint, 0.5, 50, 1)
    assert (T2.cpu() - Tref).norm() / Tref.norm() <= 1e-3


def test_cost_matrix_cosine_is_differentiable():
    """model/ot.py:11-21 is plain differentiable torch in the reference; here forward and backward are kernels."""
    _require_gpu()
    from meme_challenge_b200.model import ot
    torch.manual_seed(2)
    x = torch.randn(3, 12, 64, device=DEV, requires_grad=True)
    y = torch.randn(3, 10, 64, device=DEV, requires_grad=True)
    w = torch.randn(3, 12, 10, device=DEV)
    (ot.cost_matrix_cosine(x, y) * w).sum().backward()
    xr = x.detach().clone().requires_grad_(True)
    yr = y.detach().clone().requires_grad_(True)
    ref = 1 - torch.nn.functional.normalize(xr, p=2, dim=-1, eps=1e-5) @ torch.nn.functional.normalize(yr, p=2, dim=-1, eps=1e-5).transpose(1, 2)
    (ref * w).sum().backward()
    assert torch.allclose(ot.cost_matrix_cosine(x, y).detach(), ref.detach(), atol=1e-5)
    assert torch.allclose(x.grad, xr.grad, rtol=1e-3, atol=1e-5)
    assert torch.allclose(y.grad, yr.grad, rtol=1e-3, atol=1e-5)


def test_ot_c5_shape_and_gradient_against_oracle():
    """C5 shape: M = 64 tokens, N = 100 regions, D = 768, ragged pads; distance + d/d(emb)."""
    _require_gpu()
    from meme_challenge_b200.model import ot
    torch.manual_seed(8)
    B, M, N, D = 16, 64, 100, 768
    txt = torch.randn(B, M, D)
    img = torch.randn(B, N, D)
    tl = torch.randint(8, M + 1, (B,))
    nb = torch.randint(36, N + 1, (B,))
    txt_pad = torch.arange(M).unsqueeze(0) >= tl.unsqueeze(1)
    img_pad = torch.arange(N).unsqueeze(0) >= nb.unsqueeze(1)
    xr, yr = txt.clone().requires_grad_(True), img.clone().requires_grad_(True)
    dref, Tref, _ = O.optimal_transport_dist(xr, yr, txt_pad, img_pad)
    w = torch.randn(B)
    (dref * w).sum().backward()
    x, y = txt.to(DEV).requires_grad_(True), img.to(DEV).requires_grad_(True)
    dist, T, _ = ot.optimal_transport_plan(x, y, txt_pad.to(DEV), img_pad.to(DEV))
    (dist * w.to(DEV)).sum().backward()
    assert (T.cpu() - Tref).norm() / Tref.norm() <= 1e-3
    assert ((dist.detach().cpu() - dref.detach()).abs() / dref.detach().abs()).max() <= 1e-3
    assert (x.grad.cpu() - xr.grad).norm() / xr.grad.norm() <= 1e-3
    assert (y.grad.cpu() - yr.grad).norm() / yr.grad.norm() <= 1e-3
    # marginals of the plan: rows/cols sum to 1/len on the valid block
    s = T.sum((1, 2)).cpu()
    assert torch.allclose(s, torch.ones(B), atol=1e-3)


# ----------------------------------------------------------------------------------------------
# pretraining heads (model/pretrain.py) against golden losses from the unmodified reference
# ----------------------------------------------------------------------------------------------
def _pretrain_setup(golden_dir):
    from meme_challenge_b200.model.model import UniterConfig
    from meme_challenge_b200.model.pretrain import UniterForPretraining
    from oracle.make_golden import LABEL_DIM
    g = np.load(os.path.join(golden_dir, "tiny_pretrain.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    b = {k[3:]: torch.from_numpy(g[k]).to(DEV) for k in g.files if k.startswith("in.")}
    b["ot_inputs"] = {k[3:]: (torch.from_numpy(g[k]).to(DEV) if g[k].ndim else int(g[k])) for k in g.files if k.startswith("ot.")}
    m = UniterForPretraining(UniterConfig.from_dict(TINY), IMG_DIM, LABEL_DIM)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd, strict=True)
    return g, sd, b, m.to(DEV).eval()


MEAN_LOSS_RTOL = 1e-3


def test_vocab_cross_entropy_fused_against_torch():
    """EPI_CE_STATS / EPI_CE_GRAD (model/layer.py:204-221 decoder + model/pretrain.py:97-98 F.cross_entropy): the
    fused path never materialises the [n, 28996] logits. Against fp32 torch on the same bf16-rounded operands:
    per-row loss, d hidden, d weight (tied word embeddings), d bias."""
    _require_gpu()
    from meme_challenge_b200 import functional as F_
    torch.manual_seed(3)
    for n, V, K in ((150, 28996, 768), (7, 1000, 64), (300, 515, 128)):
        x = (torch.randn(n, K, device=DEV) * 0.7).bfloat16().requires_grad_(True)
        W = torch.nn.Parameter((torch.randn(V, K, device=DEV) * 0.05).bfloat16().float())
        bias = torch.nn.Parameter(torch.randn(V, device=DEV) * 0.1)
        tgt = torch.randint(0, V, (n,), device=DEV)
        tgt[0] = V - 1            # last (ragged) column tile
        wts = torch.rand(n, device=DEV) + 0.5
        loss = F_.vocab_cross_entropy(x, W, bias, tgt)
        (loss * wts).sum().backward()
        xr = x.detach().float().requires_grad_(True)
        Wr = W.detach().clone().requires_grad_(True)
        br = bias.detach().clone().requires_grad_(True)
        ref = torch.nn.functional.cross_entropy(xr @ Wr.t() + br, tgt, reduction="none")
        (ref * wts).sum().backward()
        assert torch.allclose(loss, ref, rtol=1e-4, atol=2e-4), (n, V, (loss - ref).abs().max().item())
        for got, want, name in ((x.grad.float(), xr.grad, "dx"), (W.grad, Wr.grad, "dW"), (bias.grad, br.grad, "db")):
            rel = ((got - want).norm() / want.norm()).item()
            assert rel < 1e-2, (n, V, name, rel)


def test_pretraining_tasks_against_reference_golden(golden_dir):
    _require_gpu()
    g, sd, b, m = _pretrain_setup(golden_dir)
    with torch.no_grad():
        # per-sample losses, relative to the largest one (measured 5e-4 / 2e-3 / 3e-4 / 6e-4 / 7e-4)
        for task, rtol in (("mlm", 3e-3), ("mrfr", 1e-2), ("itm", 2e-3), ("mrc-kl", 4e-3), ("mrc", 4e-3)):
            got = m(b, task).float().cpu().numpy()
            want = g["loss." + task]
            assert got.shape == want.shape, task
            err = np.abs(got - want).max() / (np.abs(want).max() + 1e-6)
            assert err < rtol, (task, err)
            # the quantity training optimises: the mean task loss, within MEAN_LOSS_RTOL of the reference's
            mean_err = abs(float(got.mean()) - float(want.mean())) / (abs(float(want.mean())) + 1e-12)
            if os.environ.get("B200U_PRINT_REL"):
                print("pretrain %s: max rel err %.2e, mean-loss rel err %.2e" % (task, err, mean_err))
            assert mean_err < MEAN_LOSS_RTOL, (task, mean_err)
        scores = m(b, "mlm", compute_loss=False).float().cpu().numpy()
        assert np.abs(scores - g["scores.mlm"]).max() < 2e-2
    # ITM computes the OT distance like the reference does (and drops it); check it against the oracle
    bc = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in b.items() if k != "ot_inputs"}
    bc["ot_inputs"] = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in b["ot_inputs"].items()}
    with torch.no_grad():
        _, ot_ref = O.pretrain_forward(sd, TINY, bc, "itm")
    pos, neg = m.last_ot_loss
    tgt = b["targets"].cpu()
    assert torch.allclose(pos.float().cpu(), ot_ref[tgt == 1], rtol=3e-2, atol=1e-3)
    assert torch.allclose(neg.float().cpu(), ot_ref[tgt == 0], rtol=3e-2, atol=1e-3)
    with pytest.raises(ValueError):
        m(b, "nope")


def test_pretraining_gradients_against_oracle(golden_dir):
    """MLM + MRFR backward through the tied decoder / feat_regress weights."""
    _require_gpu()
    g, sd, b, m = _pretrain_setup(golden_dir)
    bc = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in b.items() if k != "ot_inputs"}
    for task in ("mlm", "mrfr"):
        for p in m.parameters():
            p.grad = None
        m(b, task).float().mean().backward()
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        sdr["cls.predictions.decoder.weight"] = sdr["uniter.embeddings.word_embeddings.weight"]
        sdr["feat_regress.weight"] = sdr["uniter.img_embeddings.img_linear.weight"]
        O.pretrain_forward(sdr, TINY, bc, task).mean().backward()
        checked = 0
        for n, p in m.named_parameters():
            want = sdr[n].grad
            if want is None or want.abs().max() < 1e-7:
                continue
            assert p.grad is not None, (task, n)
            c = _cos(p.grad.detach().cpu(), want)
            assert c > 0.98, (task, n, c)
            checked += 1
        assert checked > 30
