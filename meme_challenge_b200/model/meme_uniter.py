"""MemeUniter (mirror of the reference model/meme_uniter.py:6-21): UniterModel -> pooler ->
Linear(H, n_classes). Logits are fp32 [B, n_classes] like the reference's."""
from torch import nn

from .. import functional as F_
from .model import UniterModel


class MemeUniter(nn.Module):

    def __init__(self, uniter_model: UniterModel, hidden_size: int, n_classes: int):
        super().__init__()
        self.uniter_model = uniter_model
        self.n_classes = n_classes
        self.linear = nn.Linear(hidden_size, n_classes)

    def forward(self, **kwargs):
        out = self.uniter_model(**kwargs)
        out = self.uniter_model.pooler(out)
        out = F_.SmallLinearFn.apply(out, self.linear.weight, self.linear.bias)
        return out
