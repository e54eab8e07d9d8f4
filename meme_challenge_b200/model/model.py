"""UNITER model (mirror of the reference model/model.py) on b200u kernels.

Public surface kept from the reference: UniterConfig (model.py:24-114), UniterPreTrainedModel with
init_weights / from_pretrained incl. the gamma/beta rename (133-214), UniterTextEmbeddings,
UniterImageEmbeddings, UniterEncoder and UniterModel.forward(input_ids, position_ids, img_feat,
img_pos_feat, attention_mask, gather_index=None, img_masks=None, output_all_encoded_layers=True,
txt_type_ids=None, img_type_ids=None) (336-367). Parameter names/shapes are identical, so
checkpoints move both ways. Differences a caller can see: encoder activations are bf16 CUDA
tensors (fp32 in the reference), and the model must live on an sm_100 GPU (no CPU fallback).
"""
import copy
import json
import logging
from io import open

import torch
from torch import nn

from .. import _lib, ops
from .. import functional as F_
from ..flat import FlatStore
from ..normalization import FusedLayerNorm
from .layer import BertLayer, BertPooler

logger = logging.getLogger(__name__)


class UniterConfig(object):
    """Configuration class to store the configuration of a `UniterModel` (model/model.py:24-114)."""

    def __init__(self, vocab_size_or_config_json_file, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072, hidden_act="gelu",
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 max_position_embeddings=512, type_vocab_size=2, initializer_range=0.02):
        if isinstance(vocab_size_or_config_json_file, str):
            with open(vocab_size_or_config_json_file, "r", encoding='utf-8') as reader:
                json_config = json.loads(reader.read())
            for key, value in json_config.items():
                self.__dict__[key] = value
        elif isinstance(vocab_size_or_config_json_file, int):
            self.vocab_size = vocab_size_or_config_json_file
            self.hidden_size = hidden_size
            self.num_hidden_layers = num_hidden_layers
            self.num_attention_heads = num_attention_heads
            self.hidden_act = hidden_act
            self.intermediate_size = intermediate_size
            self.hidden_dropout_prob = hidden_dropout_prob
            self.attention_probs_dropout_prob = attention_probs_dropout_prob
            self.max_position_embeddings = max_position_embeddings
            self.type_vocab_size = type_vocab_size
            self.initializer_range = initializer_range
        else:
            raise ValueError("First argument must be either a vocabulary size "
                             "(int) or the path to a pretrained model config "
                             "file (str)")

    @classmethod
    def from_dict(cls, json_object):
        config = UniterConfig(vocab_size_or_config_json_file=-1)
        for key, value in json_object.items():
            config.__dict__[key] = value
        return config

    @classmethod
    def from_json_file(cls, json_file):
        with open(json_file, "r", encoding='utf-8') as reader:
            text = reader.read()
        return cls.from_dict(json.loads(text))

    def __repr__(self):
        return str(self.to_json_string())

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def to_json_string(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"


class UniterPreTrainedModel(nn.Module):
    """Weight initialisation and pretrained-checkpoint loading (model/model.py:117-214)."""

    def __init__(self, config, *inputs, **kwargs):
        super().__init__()
        if not isinstance(config, UniterConfig):
            raise ValueError(
                "Parameter config in `{}(config)` should be an instance of "
                "class `UniterConfig`. To create a model from a Google "
                "pretrained model use "
                "`model = {}.from_pretrained(PRETRAINED_MODEL_NAME)`".format(
                    self.__class__.__name__, self.__class__.__name__))
        self.config = config

    def init_weights(self, module):
        """model/model.py:133-146: N(0, initializer_range) weights, unit LayerNorm, zero biases."""
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, FusedLayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    @classmethod
    def from_pretrained(cls, config_file, state_dict, *inputs, **kwargs):
        """model/model.py:148-214: build from a config json + state dict; legacy `gamma`/`beta`
        keys become `weight`/`bias`, a leading `bert.` prefix is stripped, missing / unexpected keys
        are logged and only shape errors raise."""
        config = UniterConfig.from_json_file(config_file)
        logger.info("Model config {}".format(config))
        model = cls(config, *inputs, **kwargs)
        old_keys, new_keys = [], []
        for key in state_dict.keys():
            new_key = None
            if 'gamma' in key:
                new_key = key.replace('gamma', 'weight')
            if 'beta' in key:
                new_key = key.replace('beta', 'bias')
            if new_key:
                old_keys.append(key)
                new_keys.append(new_key)
        for old_key, new_key in zip(old_keys, new_keys):
            state_dict[new_key] = state_dict.pop(old_key)

        missing_keys, unexpected_keys, error_msgs = [], [], []
        metadata = getattr(state_dict, '_metadata', None)
        state_dict = state_dict.copy()
        if metadata is not None:
            state_dict._metadata = metadata

        def load(module, prefix=''):
            local_metadata = ({} if metadata is None else metadata.get(prefix[:-1], {}))
            module._load_from_state_dict(state_dict, prefix, local_metadata, True, missing_keys,
                                         unexpected_keys, error_msgs)
            for name, child in module._modules.items():
                if child is not None:
                    load(child, prefix + name + '.')
        start_prefix = ''
        if not hasattr(model, 'bert') and any(s.startswith('bert.') for s in state_dict.keys()):
            start_prefix = 'bert.'
        load(model, prefix=start_prefix)
        if len(missing_keys) > 0:
            logger.info("Weights of {} not initialized from pretrained model: {}".format(
                model.__class__.__name__, missing_keys))
        if len(unexpected_keys) > 0:
            logger.info("Weights from pretrained model not used in {}: {}".format(
                model.__class__.__name__, unexpected_keys))
        if len(error_msgs) > 0:
            raise RuntimeError('Error(s) in loading state_dict for {}:\n\t{}'.format(
                model.__class__.__name__, "\n\t".join(error_msgs)))
        return model


class UniterTextEmbeddings(nn.Module):
    """model/model.py:217-245; forward = one fused gather+sum+LayerNorm+dropout kernel."""

    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = FusedLayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, input_ids, position_ids, token_type_ids=None, _rt=None):
        rt = _rt if _rt is not None else F_.Runtime(None, False, None, 0.0, 0.0)
        return F_.TxtEmbedFn.apply(self.LayerNorm.weight, self, input_ids, position_ids,
                                   token_type_ids, rt)


class UniterImageEmbeddings(nn.Module):
    """model/model.py:248-272; img_linear on tcgen05, everything after it in one fused kernel."""

    def __init__(self, config, img_dim):
        super().__init__()
        self.img_linear = nn.Linear(img_dim, config.hidden_size)
        self.img_layer_norm = FusedLayerNorm(config.hidden_size, eps=1e-12)
        self.pos_layer_norm = FusedLayerNorm(config.hidden_size, eps=1e-12)
        self.pos_linear = nn.Linear(7, config.hidden_size)
        self.mask_embedding = nn.Embedding(2, img_dim, padding_idx=0)
        self.LayerNorm = FusedLayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, img_feat, img_pos_feat, type_embeddings, img_masks=None, _rt=None,
                _img_type_ids=None):
        """`type_embeddings` is the token-type TABLE here (the reference passes the looked-up rows,
        model/model.py:315-318; the lookup is fused into the kernel)."""
        if _rt is None:
            raise RuntimeError("UniterImageEmbeddings.forward runs inside UniterModel")
        return F_.ImgEmbedFn.apply(self.LayerNorm.weight, self, type_embeddings, img_feat, img_pos_feat,
                                   _img_type_ids, img_masks, _rt)


class UniterEncoder(nn.Module):
    """model/model.py:275-292."""

    def __init__(self, config):
        super().__init__()
        layer = BertLayer(config)
        self.layer = nn.ModuleList([copy.deepcopy(layer) for _ in range(config.num_hidden_layers)])
        self._infer_cache = {}

    def forward(self, input_, attention_mask, output_all_encoded_layers=True, _rt=None):
        all_encoder_layers = []
        hidden_states = input_
        cb = getattr(_rt, "layer_cb", None)
        for i, layer_module in enumerate(self.layer):
            if cb is not None and hidden_states.requires_grad:
                # fires once the backward of layer i is enqueued (train.py gradient buckets)
                hidden_states.register_hook(lambda g, i=i, cb=cb: cb(i))
            hidden_states = layer_module(hidden_states, attention_mask, _rt=_rt, _layer_idx=i,
                                         _infer_cache=self._infer_cache)
            if output_all_encoded_layers:
                all_encoder_layers.append(hidden_states)
        if not output_all_encoded_layers:
            all_encoder_layers.append(hidden_states)
        return all_encoder_layers


class UniterModel(UniterPreTrainedModel):
    """Joint vision-language encoder (model/model.py:295-367)."""

    def __init__(self, config, img_dim):
        super().__init__(config)
        self.embeddings = UniterTextEmbeddings(config)
        self.img_embeddings = UniterImageEmbeddings(config, img_dim)
        self.encoder = UniterEncoder(config)
        self.pooler = BertPooler(config)
        self.apply(self.init_weights)
        self._store = FlatStore(self)
        self._seed_state = None
        self.gemm_impl = 0  # 0 = tcgen05 product path; 1 = SIMT debug kernel (bring-up only)

    # ---------------------------------------------------------------- b200u plumbing
    def flat_store(self):
        """Flat fp32 params / grads / bf16 shadow of this model (built lazily on the GPU)."""
        return self._store.ensure()

    def refresh_weights(self):
        """Force the bf16 weight shadow to be rebuilt (after editing `.data` by hand)."""
        self._store.ensure().refresh_shadow(force=True)

    def _runtime(self):
        st = self._store.ensure()
        st.refresh_shadow()
        training = self.training
        seed = None
        if torch.is_grad_enabled():
            st.attach_grads()
        if training and (self.config.hidden_dropout_prob > 0 or self.config.attention_probs_dropout_prob > 0):
            if self._seed_state is None or self._seed_state.device != st.flat.device:
                s0 = int(torch.randint(0, 2 ** 62, (1,)).item())  # honours torch.manual_seed
                self._seed_state = torch.tensor([s0], device=st.flat.device, dtype=torch.int64)
            seed = self._seed_state.clone()
            ops.counter_add(self._seed_state, 0x9E3779B97F4A7C15)
        rt = F_.Runtime(st, training, seed, self.config.hidden_dropout_prob,
                        self.config.attention_probs_dropout_prob, self.gemm_impl)
        rt.layer_cb = getattr(self, "_layer_grad_ready_cb", None)
        rt.sparse_word_cb = getattr(self, "_sparse_word_cb", None)
        return rt

    # ---------------------------------------------------------------- reference API
    def _compute_txt_embeddings(self, input_ids, position_ids, txt_type_ids=None, _rt=None):
        return self.embeddings(input_ids, position_ids, txt_type_ids, _rt=_rt or self._runtime())

    def _compute_img_embeddings(self, img_feat, img_pos_feat, img_masks=None, img_type_ids=None, _rt=None):
        return self.img_embeddings(img_feat, img_pos_feat, self.embeddings.token_type_embeddings.weight,
                                   img_masks, _rt=_rt or self._runtime(), _img_type_ids=img_type_ids)

    def _compute_img_txt_embeddings(self, input_ids, position_ids, img_feat, img_pos_feat, gather_index,
                                    img_masks=None, txt_type_ids=None, img_type_ids=None, _rt=None):
        rt = _rt or self._runtime()
        txt_emb = self._compute_txt_embeddings(input_ids, position_ids, txt_type_ids, _rt=rt)
        img_emb = self._compute_img_embeddings(img_feat, img_pos_feat, img_masks, img_type_ids, _rt=rt)
        # align back to most compact input (model/model.py:329-333), bit-exact row gather
        return F_.GatherFn.apply(txt_emb, img_emb, gather_index)

    def forward(self, input_ids, position_ids, img_feat, img_pos_feat, attention_mask=None,
                gather_index=None, img_masks=None, output_all_encoded_layers=True,
                txt_type_ids=None, img_type_ids=None, attn_masks=None):
        if attention_mask is None:
            attention_mask = attn_masks  # batch-dict key used by the pretraining collates
        if attention_mask is None:
            raise TypeError("forward() missing required argument: 'attention_mask'")
        rt = self._runtime()
        # compute self-attention mask (model/model.py:342-345); kept as one [B, L] row per sample
        extended_attention_mask = (1.0 - attention_mask.to(dtype=torch.float32)) * -10000.0
        extended_attention_mask = extended_attention_mask.contiguous()

        if input_ids is None:
            embedding_output = self._compute_img_embeddings(img_feat, img_pos_feat, img_masks,
                                                            img_type_ids, _rt=rt)
        elif img_feat is None:
            embedding_output = self._compute_txt_embeddings(input_ids, position_ids, txt_type_ids, _rt=rt)
        else:
            embedding_output = self._compute_img_txt_embeddings(
                input_ids, position_ids, img_feat, img_pos_feat, gather_index, img_masks,
                txt_type_ids, img_type_ids, _rt=rt)

        encoded_layers = self.encoder(embedding_output, extended_attention_mask,
                                      output_all_encoded_layers=output_all_encoded_layers, _rt=rt)
        if not output_all_encoded_layers:
            encoded_layers = encoded_layers[-1]
        return encoded_layers
