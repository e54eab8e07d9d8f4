"""Wasserstein distance / optimal transport (mirror of the reference model/ot.py) on b200u kernels.

Same functions, argument order and result shapes: cost_matrix_cosine -> [B, Lx, Ly]; ipot ->
T [B, N, M] (image-major, ot.py:36-66); optimal_transport_dist -> [B] with the gradient flowing
only through the cost (T detached, ot.py:82-84). fp32 throughout (model/pretrain.py:188-190).
The IPOT loop is one kernel launch instead of ~350.
"""
import ctypes as C

import torch

from .. import _lib, ops

P = _lib.ptr


def _u8(pad):
    return pad.to(torch.uint8).contiguous()


def _cost_fwd(x, y, x_pad, y_pad, eps):
    B, M, D = x.shape
    N = y.shape[1]
    cost = torch.empty(B, M, N, device=x.device, dtype=torch.float32)
    xinv = torch.empty(B, M, device=x.device, dtype=torch.float32)
    yinv = torch.empty(B, N, device=x.device, dtype=torch.float32)
    ops._call("b200u_cosine_cost", P(x), P(y), P(x_pad), P(y_pad), P(cost), P(xinv), P(yinv), B, M, N, D,
              float(eps))
    return cost, xinv, yinv


class _CosineCost(torch.autograd.Function):
    """cosine distance with its gradient: the backward kernel of the OT distance takes the upstream weights of the
    cost entries as a [B, N, M] matrix (there the transport plan), here d loss / d cost transposed."""

    @staticmethod
    def forward(ctx, x, y, eps):
        xf = x.detach().float().contiguous()
        yf = y.detach().float().contiguous()
        cost, xinv, yinv = _cost_fwd(xf, yf, None, None, eps)
        ctx.save_for_backward(xf, yf, xinv, yinv)
        ctx.in_dtypes = (x.dtype, y.dtype)
        return cost

    @staticmethod
    def backward(ctx, dcost):
        x, y, xinv, yinv = ctx.saved_tensors
        B, M, D = x.shape
        N = y.shape[1]
        w = dcost.float().transpose(1, 2).contiguous()
        one = torch.ones(B, device=x.device, dtype=torch.float32)
        dx = torch.empty_like(x)
        dy = torch.empty_like(y)
        ops._call("b200u_cosine_cost_bwd", P(x), P(y), P(xinv), P(yinv), None, None, P(w), P(one), P(dx), P(dy),
                  B, M, N, D)
        return dx.to(ctx.in_dtypes[0]), dy.to(ctx.in_dtypes[1]), None


def cost_matrix_cosine(x, y, eps=1e-5):
    """[B, L_x, D] [B, L_y, D] -> [B, Lx, Ly] cosine distance (ot.py:11-21), differentiable like the reference's."""
    assert x.dim() == y.dim()
    assert x.size(0) == y.size(0)
    assert x.size(2) == y.size(2)
    return _CosineCost.apply(x, y, eps)


def trace(x):
    """Batched trace (ot.py:24-32)."""
    b, m, n = x.size()
    assert m == n
    return torch.diagonal(x, dim1=-2, dim2=-1).sum(-1)


@torch.no_grad()
def ipot(C_, x_len, x_pad, y_len, y_pad, joint_pad, beta, iteration, k):
    """[B, M, N], [B], [B, M], [B], [B, N], [B, M, N] -> T [B, N, M] (ot.py:35-66). x_len / y_len /
    joint_pad are implied by the pads (ot.py:74-80) and recomputed inside the kernel."""
    B, M, N = C_.shape
    cost = C_.detach().float().contiguous()
    xp, yp = _u8(x_pad), _u8(y_pad)
    T = torch.empty(B, N, M, device=cost.device, dtype=torch.float32)
    ops._call("b200u_ipot", P(cost), P(xp), P(yp), P(T), B, M, N, float(beta), int(iteration), int(k))
    return T


class _OTDist(torch.autograd.Function):
    @staticmethod
    def forward(ctx, txt_emb, img_emb, txt_pad, img_pad, beta, iteration, k):
        x = txt_emb.detach().float().contiguous()
        y = img_emb.detach().float().contiguous()
        xp, yp = _u8(txt_pad), _u8(img_pad)
        B, M, D = x.shape
        N = y.shape[1]
        cost, xinv, yinv = _cost_fwd(x, y, xp, yp, 1e-5)
        T = torch.empty(B, N, M, device=x.device, dtype=torch.float32)
        ops._call("b200u_ipot", P(cost), P(xp), P(yp), P(T), B, M, N, float(beta), int(iteration), int(k))
        dist = torch.empty(B, device=x.device, dtype=torch.float32)
        ops._call("b200u_ot_distance", P(cost), P(T), P(dist), B, M, N)
        ctx.save_for_backward(x, y, xinv, yinv, xp, yp, T)
        ctx.in_dtypes = (txt_emb.dtype, img_emb.dtype)
        ctx.mark_non_differentiable(T, cost)
        return dist, T, cost

    @staticmethod
    def backward(ctx, ddist, _dT, _dcost):
        x, y, xinv, yinv, xp, yp, T = ctx.saved_tensors
        B, M, D = x.shape
        N = y.shape[1]
        ddist = ddist.contiguous().float()
        dx = torch.empty_like(x)
        dy = torch.empty_like(y)
        ops._call("b200u_cosine_cost_bwd", P(x), P(y), P(xinv), P(yinv), P(xp), P(yp), P(T), P(ddist), P(dx),
                  P(dy), B, M, N, D)
        return dx.to(ctx.in_dtypes[0]), dy.to(ctx.in_dtypes[1]), None, None, None, None, None


def optimal_transport_dist(txt_emb, img_emb, txt_pad, img_pad, beta=0.5, iteration=50, k=1):
    """[B, M, D], [B, N, D], [B, M], [B, N] -> distance [B] (ot.py:69-85)."""
    dist, _, _ = _OTDist.apply(txt_emb, img_emb, txt_pad, img_pad, beta, iteration, k)
    return dist


def optimal_transport_plan(txt_emb, img_emb, txt_pad, img_pad, beta=0.5, iteration=50, k=1):
    """Same computation, also returning the transport plan T [B, N, M] and the masked cost."""
    return _OTDist.apply(txt_emb, img_emb, txt_pad, img_pad, beta, iteration, k)
