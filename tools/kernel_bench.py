"""In-graph timing (20 back-to-back launches, CUDA events) of the non-GEMM kernels at the C2 shape."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from meme_challenge_b200 import _lib, ops

dev = "cuda"
B, L, heads, H, I = int(os.environ.get("KB_B", "16")), 164, 12, 768, 3072
M = B * L
seed = torch.tensor([7], device=dev, dtype=torch.int64)


def timeit(name, fn, reps=20, bytes_moved=None):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    extra = "  %.0f GB/s" % (bytes_moved / us / 1e3) if bytes_moved else ""
    print("%-34s %8.2f us%s" % (name, us, extra), flush=True)


x = torch.randn(M, H, device=dev).bfloat16()
dy = torch.randn(M, H, device=dev).bfloat16()
gamma = torch.ones(H, device=dev); beta = torch.zeros(H, device=dev)
y, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-12)
dg, db_, dbias = (torch.zeros(H, device=dev) for _ in range(3))
timeit("layernorm_fwd", lambda: ops.layernorm_fwd(x, gamma, beta, 1e-12), bytes_moved=2 * M * H * 2)
timeit("layernorm_bwd (dx, dz, dbias, p=0.1)", lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg, db_, dz=True, dbias=dbias, drop=_lib.dropout_t(seed, 3, 0.1)), bytes_moved=4 * M * H * 2)
timeit("layernorm_bwd (dx only, p=0)", lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg, db_), bytes_moved=3 * M * H * 2)
big = torch.randn(M, I, device=dev).bfloat16(); out_i = torch.zeros(I, device=dev)
q3 = torch.randn(M, 3 * H, device=dev).bfloat16(); out_q = torch.zeros(3 * H, device=dev)
timeit("colsum [M,3072]", lambda: ops.colsum_accum(big, out_i), bytes_moved=M * I * 2)
timeit("colsum [M,2304]", lambda: ops.colsum_accum(q3, out_q), bytes_moved=M * 3 * H * 2)
qkv = (torch.randn(M, 3 * H, device=dev) * 0.5).bfloat16()
mask = torch.zeros(B, L, device=dev)
d = _lib.dropout_t(seed, 5, 0.1)
ctx, lse = ops.attention_fwd(qkv, mask, B, L, heads, H, drop=d)
dctx = torch.randn(M, H, device=dev).bfloat16()
timeit("attention_fwd (p=0.1)", lambda: ops.attention_fwd(qkv, mask, B, L, heads, H, drop=d), bytes_moved=4 * M * H * 2)
timeit("attention_fwd (p=0)", lambda: ops.attention_fwd(qkv, mask, B, L, heads, H), bytes_moved=4 * M * H * 2)
timeit("attention_bwd (p=0.1, 2 kernels)", lambda: ops.attention_bwd(qkv, mask, ctx, dctx, lse, B, L, heads, H, drop=d), bytes_moved=8 * M * H * 2)
timeit("attention_bwd (p=0)", lambda: ops.attention_bwd(qkv, mask, ctx, dctx, lse, B, L, heads, H), bytes_moved=8 * M * H * 2)
