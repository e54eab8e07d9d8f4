"""BERT layers of UNITER (mirror of the reference model/layer.py) on b200u kernels.

Same classes, constructor arguments, sub-module names and therefore the same state_dict keys as
the reference. BertLayer.forward runs the whole layer (model/layer.py:166-170) as one fused C call:
tcgen05 GEMMs with bias/GELU/dropout/residual epilogues, shared-memory softmax attention and
LayerNorm kernels. Activations are bf16, accumulation / LayerNorm / softmax are fp32.
"""
import math

import torch
from torch import nn

from .. import functional as F_
from ..normalization import FusedLayerNorm as BertLayerNorm


def gelu(x):
    """model/layer.py:31-37 (exact erf form); host-visible helper, the kernels fuse their own."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def swish(x):
    return x * torch.sigmoid(x)


ACT2FN = {"gelu": gelu, "relu": torch.nn.functional.relu, "swish": swish}


class GELU(nn.Module):
    def forward(self, input_):
        return gelu(input_)


class BertSelfAttention(nn.Module):
    def __init__(self, config):
        super(BertSelfAttention, self).__init__()
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError(
                "The hidden size (%d) is not a multiple of the number of attention "
                "heads (%d)" % (config.hidden_size, config.num_attention_heads))
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = int(config.hidden_size / config.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        if self.attention_head_size != 64:
            raise ValueError("b200u attention kernels are built for head size 64 "
                             "(uniter-base / uniter-large); got %d" % self.attention_head_size)
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)


class BertSelfOutput(nn.Module):
    def __init__(self, config):
        super(BertSelfOutput, self).__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class BertAttention(nn.Module):
    def __init__(self, config):
        super(BertAttention, self).__init__()
        self.self = BertSelfAttention(config)
        self.output = BertSelfOutput(config)


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super(BertIntermediate, self).__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        if config.hidden_act != "gelu":
            raise ValueError("b200u fuses the exact-erf GELU (config hidden_act='gelu'); got %r"
                             % (config.hidden_act,))
        self.intermediate_act_fn = ACT2FN[config.hidden_act]


class BertOutput(nn.Module):
    def __init__(self, config):
        super(BertOutput, self).__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class BertLayer(nn.Module):
    def __init__(self, config):
        super(BertLayer, self).__init__()
        self.attention = BertAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)

    def forward(self, hidden_states, attention_mask, _rt=None, _layer_idx=0, _infer_cache=None):
        """hidden_states bf16 [B,L,H]; attention_mask = the additive mask of model/model.py:342-345
        ([B,1,1,L] or [B,L], 0 / -10000). `_rt` is supplied by UniterEncoder."""
        if _rt is None:
            raise RuntimeError("BertLayer.forward runs inside UniterModel (it needs the model's flat "
                               "weight store); call UniterModel / UniterEncoder instead")
        mask_add = attention_mask.reshape(hidden_states.shape[0], -1)
        if torch.is_grad_enabled():
            return F_.BertLayerFn.apply(hidden_states, mask_add, self.output.LayerNorm.weight, self,
                                        _layer_idx, _rt)
        return F_.bert_layer_infer(hidden_states, mask_add, self, _layer_idx, _rt, _infer_cache)


class BertPooler(nn.Module):
    def __init__(self, config):
        super(BertPooler, self).__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        """model/layer.py:179-185: tanh(dense(hidden_states[:, 0])) -> fp32 [B, H]."""
        if hidden_states.dtype != torch.bfloat16:
            hidden_states = hidden_states.to(torch.bfloat16)
        return F_.PoolerFn.apply(hidden_states.contiguous(), self.dense.weight, self.dense.bias)


class BertPredictionHeadTransform(nn.Module):
    """model/layer.py:188-201: dense -> gelu -> LayerNorm (GELU fused into the GEMM epilogue)."""

    def __init__(self, config):
        super(BertPredictionHeadTransform, self).__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.transform_act_fn = ACT2FN[config.hidden_act] if isinstance(config.hidden_act, str) else config.hidden_act
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=1e-12)

    def forward(self, hidden_states):
        hidden_states = F_.linear(hidden_states, self.dense.weight, self.dense.bias, gelu=True)
        return self.LayerNorm(hidden_states)


class BertLMPredictionHead(nn.Module):
    """model/layer.py:204-221: decoder weight tied to the word embeddings + output-only bias;
    logits are fp32 [n_masked, vocab]."""

    def __init__(self, config, bert_model_embedding_weights):
        super(BertLMPredictionHead, self).__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(bert_model_embedding_weights.size(1), bert_model_embedding_weights.size(0),
                                 bias=False)
        self.decoder.weight = bert_model_embedding_weights
        self.bias = nn.Parameter(torch.zeros(bert_model_embedding_weights.size(0)))

    def forward(self, hidden_states):
        hidden_states = self.transform(hidden_states)
        return F_.linear(hidden_states, self.decoder.weight, self.bias, out_f32=True)

    def cross_entropy(self, hidden_states, targets):
        """F.cross_entropy(self(hidden_states), targets, reduction='none') with the decoder GEMM and the loss fused:
        the [n, vocab] logits are never written (functional.VocabCrossEntropyFn)."""
        hidden_states = self.transform(hidden_states)
        return F_.vocab_cross_entropy(hidden_states, self.decoder.weight, self.bias, targets)


class BertOnlyMLMHead(nn.Module):
    """model/layer.py:224-233."""

    def __init__(self, config, bert_model_embedding_weights):
        super(BertOnlyMLMHead, self).__init__()
        self.predictions = BertLMPredictionHead(config, bert_model_embedding_weights)

    def forward(self, sequence_output):
        return self.predictions(sequence_output)

    def cross_entropy(self, sequence_output, targets):
        return self.predictions.cross_entropy(sequence_output, targets)
