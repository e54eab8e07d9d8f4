"""In-graph timing (20 back-to-back launches, CUDA events) of the GEMM shapes of one BertLayer at the C2
shape, old vs new epilogues: fused LayerNorm vs GEMM + standalone LayerNorm, saved-gelu' multiply vs DGELU."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from meme_challenge_b200 import _lib, ops

dev = "cuda"
E = _lib
M, H, I = int(os.environ.get("GB_M", "2624")), 768, 3072
if len(sys.argv) > 1 and sys.argv[1] == "large":
    H, I = 1024, 4096
seed = torch.tensor([7], device=dev, dtype=torch.int64)
drop = _lib.dropout_t(seed, 3, 0.1)


def timeit(name, fn, flops, reps=20):
    fn(0); fn(1); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i % 2)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    print("%-44s %8.2f us  %7.1f TFLOP/s" % (name, us, flops / us / 1e6), flush=True)
    return us


def mk(m, n, k, am=0, bm=0):
    return [(torch.randn((k, m) if am else (m, k), device=dev).bfloat16(),
             (torch.randn((k, n) if bm else (n, k), device=dev) * 0.05).bfloat16()) for _ in range(2)]


gam, bet = torch.ones(H, device=dev), torch.zeros(H, device=dev)
mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
bias_h, bias_i = torch.randn(H, device=dev), torch.randn(I, device=dev)
res_h = torch.randn(M, H, device=dev).bfloat16()
res_i = torch.rand(M, I, device=dev).bfloat16()
out_h, out_h2 = torch.empty(M, H, device=dev).bfloat16(), torch.empty(M, H, device=dev).bfloat16()
out_i, out_i2 = torch.empty(M, I, device=dev).bfloat16(), torch.empty(M, I, device=dev).bfloat16()
cs = torch.zeros(I, device=dev)

for name, K in (("attn_out", H), ("ffn2", I)):
    ab = mk(M, H, K)
    fl = 2.0 * M * H * K
    t0 = timeit(name + "_fwd BIAS_DROP_RES", lambda i: ops.gemm(*ab[i], bias=bias_h, res=res_h, drop=drop, epilogue=E.EPI_BIAS_DROP_RES, out=out_h), fl)
    t1 = timeit("  + layernorm_fwd (separate launch)", lambda i: (ops.gemm(*ab[i], bias=bias_h, res=res_h, drop=drop, epilogue=E.EPI_BIAS_DROP_RES, out=out_h), ops.layernorm_fwd(out_h, gam, bet, 1e-12)), fl)
    t2 = timeit(name + "_fwd BIAS_DROP_RES_LN (fused)", lambda i: ops.gemm(*ab[i], bias=bias_h, res=res_h, drop=drop, epilogue=E.EPI_BIAS_DROP_RES_LN, out=out_h, out2=out_h2, ln=(gam, bet, 1e-12, mean, rstd)), fl)
ab = mk(M, I, H)
fl = 2.0 * M * I * H
timeit("ffn1_fwd BIAS_GELU", lambda i: ops.gemm(*ab[i], bias=bias_i, epilogue=E.EPI_BIAS_GELU, out=out_i, out2=out_i2), fl)
timeit("ffn1_fwd BIAS_GELU_DG", lambda i: ops.gemm(*ab[i], bias=bias_i, epilogue=E.EPI_BIAS_GELU_DG, out=out_i, out2=out_i2), fl)
timeit("ffn1_fwd STORE (no gelu)", lambda i: ops.gemm(*ab[i], bias=bias_i, epilogue=E.EPI_STORE, out=out_i), fl)
ab = mk(M, I, H, 0, 1)
timeit("ffn2_dgrad DGELU", lambda i: ops.gemm(*ab[i], b_mn=True, res=res_i, epilogue=E.EPI_DGELU, out=out_i), fl)
timeit("  + colsum (separate launch)", lambda i: (ops.gemm(*ab[i], b_mn=True, res=res_i, epilogue=E.EPI_DGELU, out=out_i), ops.colsum_accum(out_i, cs)), fl)
timeit("ffn2_dgrad MUL", lambda i: ops.gemm(*ab[i], b_mn=True, res=res_i, epilogue=E.EPI_MUL, out=out_i), fl)
timeit("ffn2_dgrad MUL + colsum (fused)", lambda i: ops.gemm(*ab[i], b_mn=True, res=res_i, epilogue=E.EPI_MUL, out=out_i, colsum=cs), fl)
ab = mk(M, 3 * H, H)
timeit("qkv_fwd STORE bn=256", lambda i: ops.gemm(*ab[i], bias=torch.zeros(3 * H, device=dev), epilogue=E.EPI_STORE, block_n=256), 2.0 * M * 3 * H * H)
timeit("qkv_fwd STORE bn=128", lambda i: ops.gemm(*ab[i], bias=torch.zeros(3 * H, device=dev), epilogue=E.EPI_STORE, block_n=128), 2.0 * M * 3 * H * H)
# remaining shapes of the layer (backward)
ab = mk(M, H, I, 0, 1)
timeit("ffn1_dgrad ADD", lambda i: ops.gemm(*ab[i], b_mn=True, res=res_h, epilogue=E.EPI_ADD, out=out_h), 2.0 * M * H * I)
ab = mk(M, H, H, 0, 1)
timeit("attn_out_dgrad STORE", lambda i: ops.gemm(*ab[i], b_mn=True, epilogue=E.EPI_STORE, out=out_h), 2.0 * M * H * H)
ab = mk(M, H, 3 * H, 0, 1)
timeit("qkv_dgrad ADD", lambda i: ops.gemm(*ab[i], b_mn=True, res=res_h, epilogue=E.EPI_ADD, out=out_h), 2.0 * M * H * 3 * H)
for nm, (m, n) in (("ffn2_wgrad", (H, I)), ("ffn1_wgrad", (I, H)), ("attn_out_wgrad", (H, H)), ("qkv_wgrad", (3 * H, H))):
    ab = mk(m, n, M, 1, 1)
    acc = torch.zeros(m, n, device=dev)
    timeit(nm + " ATOMIC_F32", lambda i: ops.gemm(*ab[i], a_mn=True, b_mn=True, epilogue=E.EPI_ATOMIC_F32, out=acc), 2.0 * M * m * n)
