"""FusedLayerNorm — drop-in for apex.normalization.fused_layer_norm.FusedLayerNorm as the
reference uses it (model/model.py:16,229,252-258; model/layer.py:25,108,149; model/pretrain.py:12):
y = (x - mean) / sqrt(var + eps) * weight + bias over the last dimension, biased variance.
Backed by b200u_layernorm_{fwd,bwd}; CUDA tensors only (bf16 or fp32)."""
import torch
from torch import nn

from . import functional as F_


class FusedLayerNorm(nn.Module):
    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True):
        super().__init__()
        if isinstance(normalized_shape, int):
            normalized_shape = (normalized_shape,)
        if len(normalized_shape) != 1:
            raise ValueError("b200u FusedLayerNorm normalises over the last dimension only")
        if not elementwise_affine:
            raise ValueError("b200u FusedLayerNorm is always affine (as every call site in the reference)")
        self.normalized_shape = torch.Size(normalized_shape)
        self.eps = eps
        self.elementwise_affine = True
        self.weight = nn.Parameter(torch.ones(*normalized_shape))
        self.bias = nn.Parameter(torch.zeros(*normalized_shape))

    def forward(self, x):
        return F_.LayerNormFn.apply(x, self.weight, self.bias, self.eps)

    def extra_repr(self):
        return "{}, eps={}".format(tuple(self.normalized_shape), self.eps)
