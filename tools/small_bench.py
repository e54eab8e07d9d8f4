"""In-graph times of the small kernels at both ends of a fused-window UNITER-base step (B = 32 memes, T = 64,
R = 100, L = 164, H = 768): 20 back-to-back launches in a CUDA graph, warm L2, alternating operand sets.
    python tools/small_bench.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from meme_challenge_b200 import _lib, ops  # noqa: E402
from meme_challenge_b200.roofline import _time_graph  # noqa: E402

P = _lib.ptr
dev = torch.device("cuda", 0)
B, T, R, L, H, D = 32, 64, 100, 164, 768, 2048
torch.manual_seed(0)
seed = torch.tensor([7], device=dev, dtype=torch.int64)


def two(f):
    return [f(), f()]


res = {}

# ---- embedding scatter-add: word ids (mostly distinct), position ids (repeat down the batch)
d = two(lambda: torch.randn(B * T, H, device=dev).bfloat16())
word_ids = torch.randint(1000, 28000, (B, T), device=dev)
word_ids[:, 0] = 101
word_ids[:, 40:] = 0
pos_ids = torch.arange(T, device=dev).unsqueeze(0).repeat(B, 1).contiguous()
table = torch.zeros(28996, H, device=dev)
ptab = torch.zeros(512, H, device=dev)
res["scatter_add_word"] = _time_graph(lambda i: ops._call(
    "b200u_embedding_scatter_add", P(d[i]), P(word_ids), T, T, C.c_longlong(0), P(table), B * T, H, C.c_longlong(0), C.c_longlong(28996)))
res["scatter_add_pos"] = _time_graph(lambda i: ops._call(
    "b200u_embedding_scatter_add", P(d[i]), P(pos_ids), T, T, C.c_longlong(0), P(ptab), B * T, H, C.c_longlong(-1), C.c_longlong(512)))

# ---- image embedder forward
n = B * R
a = two(lambda: torch.randn(n, H, device=dev))
pos7 = torch.rand(n, 7, device=dev)
Wp, bp = torch.randn(H, 7, device=dev) * 0.1, torch.zeros(H, device=dev)
ty = torch.randn(2, H, device=dev) * 0.02
ones, zeros = torch.ones(H, device=dev), torch.zeros(H, device=dev)
out = torch.empty(n, H, device=dev, dtype=torch.bfloat16)
p_out, s_out = torch.empty(n, H, device=dev), torch.empty(n, H, device=dev)
stats = torch.empty(6, n, device=dev)
drop = _lib.dropout_t(seed, 2, 0.1)
res["img_embed_fwd"] = _time_graph(lambda i: ops._call(
    "b200u_img_embed_fwd", P(a[i]), P(pos7), P(Wp), P(bp), None, P(ty), P(ones), P(zeros), P(ones), P(zeros), P(ones),
    P(zeros), P(out), P(p_out), P(s_out), P(stats), n, H, 2, 1e-12, C.byref(drop)))
dp = two(lambda: torch.randn(n, H, device=dev).bfloat16())
dWp = torch.zeros(H, 7, device=dev)
res["pos_linear_wgrad"] = _time_graph(lambda i: ops._call("b200u_pos_linear_wgrad", P(dp[i]), P(pos7), P(dWp), n, H))

# ---- pooler + classification head
hid = two(lambda: torch.randn(B, L, H, device=dev).bfloat16())
W = torch.randn(H, H, device=dev) * 0.02
bias = torch.zeros(H, device=dev)
pooled = torch.empty(B, H, device=dev)
res["pooler_fwd"] = _time_graph(lambda i: ops._call("b200u_pooler_fwd", P(hid[i]), C.c_longlong(L * H), P(W), P(bias),
                                                    P(pooled), B, H))
dpooled = torch.randn(B, H, device=dev)
dW, db = torch.zeros(H, H, device=dev), torch.zeros(H, device=dev)
dh = two(lambda: torch.zeros(B, L, H, device=dev, dtype=torch.bfloat16))
res["pooler_bwd"] = _time_graph(lambda i: ops._call("b200u_pooler_bwd", P(dpooled), P(pooled), P(hid[i]),
                                                    C.c_longlong(L * H), P(W), P(dW), P(db), P(dh[i]),
                                                    C.c_longlong(L * H), B, H))
Wc = torch.randn(1, H, device=dev) * 0.02
dout = torch.randn(B, 1, device=dev)
x = torch.randn(B, H, device=dev)
dx = torch.empty(B, H, device=dev)
gw, gb = torch.zeros(1, H, device=dev), torch.zeros(1, device=dev)
res["linear_small_bwd"] = _time_graph(lambda i: ops._call("b200u_linear_small_bwd", P(dout), P(x), P(Wc), P(dx), P(gw),
                                                          P(gb), B, 1, H))

# ---- fp32-input LayerNorm backward of the embedders, gather backward
xf = two(lambda: torch.randn(n, H, device=dev))
dy = two(lambda: torch.randn(n, H, device=dev).bfloat16())
_, mean, rstd = ops.layernorm_fwd(xf[0], ones, zeros, 1e-12)
dg, dbt, dbi = (torch.zeros(H, device=dev) for _ in range(3))
res["layernorm_bwd_f32"] = _time_graph(lambda i: ops.layernorm_bwd(dy[i], xf[i], mean, rstd, ones, dg, dbt, dbias=dbi))
gi = torch.arange(L, device=dev).unsqueeze(0).repeat(B, 1).contiguous()
dj = two(lambda: torch.randn(B, L, H, device=dev).bfloat16())
res["gather_rows_bwd"] = _time_graph(lambda i: ops.gather_rows_bwd(dj[i], gi, T, R))
f32 = two(lambda: torch.randn(B * R, D, device=dev))
b16 = torch.empty(B * R, D, device=dev, dtype=torch.bfloat16)
res["cast_f32_bf16"] = _time_graph(lambda i: ops.cast_f32_to_bf16(f32[i], b16))
for k, v in res.items():
    print("%-22s %7.2f us" % (k, v))
