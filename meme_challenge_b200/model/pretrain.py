"""UNITER for pretraining (mirror of the reference model/pretrain.py) on b200u kernels.

Same classes, state_dict keys (uniter.*, cls.predictions.*, feat_regress.*, region_classifier.*,
itm_output.*; tied decoder / feat_regress weights) and `forward(batch, task, compute_loss)` task
dispatch. Encoder, head GEMMs, LayerNorms, pooler and the IPOT alignment run on b200u kernels; the
MLM loss is the fused vocabulary GEMM + online log-softmax + NLL (functional.VocabCrossEntropyFn: the
[n_masked, 28996] logits are never written); the other per-task loss reductions (MSE / KL / 2-way cross
entropy over a few hundred rows) stay torch functional calls on the fp32 head outputs.
"""
from collections import defaultdict

import torch
from torch import nn
from torch.nn import functional as F

from .. import functional as F_
from ..normalization import FusedLayerNorm as LayerNorm
from .layer import GELU, BertOnlyMLMHead
from .model import UniterModel, UniterPreTrainedModel
from .ot import optimal_transport_dist


class RegionFeatureRegression(nn.Module):
    " for MRM (model/pretrain.py:19-33)"

    def __init__(self, hidden_size, feat_dim, img_linear_weight):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(hidden_size, hidden_size), GELU(), LayerNorm(hidden_size, eps=1e-12))
        self.weight = img_linear_weight
        self.bias = nn.Parameter(torch.zeros(feat_dim))

    def forward(self, input_):
        hidden = F_.linear(input_, self.net[0].weight, self.net[0].bias, gelu=True)
        hidden = self.net[2](hidden)
        # F.linear(hidden, self.weight.t(), self.bias): the tied img_linear weight used transposed
        return F_.linear(hidden, self.weight, self.bias, transposed=True, out_f32=True)


class RegionClassification(nn.Module):
    " for MRC(-kl) (model/pretrain.py:36-47)"

    def __init__(self, hidden_size, label_dim):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(hidden_size, hidden_size), GELU(), LayerNorm(hidden_size, eps=1e-12),
                                 nn.Linear(hidden_size, label_dim))

    def forward(self, input_):
        hidden = F_.linear(input_, self.net[0].weight, self.net[0].bias, gelu=True)
        hidden = self.net[2](hidden)
        return F_.linear(hidden, self.net[3].weight, self.net[3].bias, out_f32=True)


class UniterForPretraining(UniterPreTrainedModel):
    """ UNITER pretraining (model/pretrain.py:50-233) """

    def __init__(self, config, img_dim, img_label_dim):
        super().__init__(config)
        self.uniter = UniterModel(config, img_dim)
        self.cls = BertOnlyMLMHead(config, self.uniter.embeddings.word_embeddings.weight)
        self.feat_regress = RegionFeatureRegression(config.hidden_size, img_dim,
                                                    self.uniter.img_embeddings.img_linear.weight)
        self.region_classifier = RegionClassification(config.hidden_size, img_label_dim)
        self.itm_output = nn.Linear(config.hidden_size, 2)
        # MLM loss through the fused vocabulary GEMM + cross entropy (False: materialise fp32 logits like the reference)
        self.fused_vocab_ce = True
        self.apply(self.init_weights)

    def forward(self, batch, task, compute_loss=True):
        batch = defaultdict(lambda: None, batch)
        input_ids = batch['input_ids']
        position_ids = batch['position_ids']
        img_feat = batch['img_feat']
        img_pos_feat = batch['img_pos_feat']
        attention_mask = batch['attn_masks']
        gather_index = batch['gather_index']
        if task == 'mlm':
            return self.forward_mlm(input_ids, position_ids, img_feat, img_pos_feat, attention_mask,
                                    gather_index, batch['txt_labels'], compute_loss)
        elif task == 'mrfr':
            return self.forward_mrfr(input_ids, position_ids, img_feat, img_pos_feat, attention_mask,
                                     gather_index, batch['img_masks'], batch['img_mask_tgt'],
                                     batch['feat_targets'], compute_loss)
        elif task == 'itm':
            return self.forward_itm(input_ids, position_ids, img_feat, img_pos_feat, attention_mask,
                                    gather_index, batch['targets'], batch['ot_inputs'], compute_loss)
        elif task.startswith('mrc'):
            return self.forward_mrc(input_ids, position_ids, img_feat, img_pos_feat, attention_mask,
                                    gather_index, batch['img_masks'], batch['img_mask_tgt'],
                                    batch['label_targets'], task, compute_loss)
        else:
            raise ValueError('invalid task')

    def forward_mlm(self, input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                    txt_labels, compute_loss=True):
        sequence_output = self.uniter(input_ids, position_ids, img_feat, img_pos_feat, attention_mask,
                                      gather_index, output_all_encoded_layers=False)
        sequence_output = sequence_output[:, :input_ids.size(1), :]       # text part only
        masked_output = self._compute_masked_hidden(sequence_output, txt_labels != -1)
        if compute_loss:
            if self.fused_vocab_ce:
                # decoder GEMM + log-softmax + NLL in one pass over the vocabulary: no [n_masked, 28996] logits
                return self.cls.cross_entropy(masked_output, txt_labels[txt_labels != -1])
            return F.cross_entropy(self.cls(masked_output), txt_labels[txt_labels != -1], reduction='none')
        return self.cls(masked_output)

    def _compute_masked_hidden(self, hidden, mask):
        """ get only the masked region (don't compute unnecessary hiddens) """
        mask = mask.unsqueeze(-1).expand_as(hidden)
        return hidden[mask].contiguous().view(-1, hidden.size(-1))

    def forward_mrfr(self, input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                     img_masks, img_mask_tgt, feat_targets, compute_loss=True):
        sequence_output = self.uniter(input_ids, position_ids, img_feat, img_pos_feat, attention_mask,
                                      gather_index, output_all_encoded_layers=False, img_masks=img_masks)
        masked_output = self._compute_masked_hidden(sequence_output, img_mask_tgt)
        prediction_feat = self.feat_regress(masked_output)
        if compute_loss:
            return F.mse_loss(prediction_feat, feat_targets, reduction='none')
        return prediction_feat

    def forward_itm(self, input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                    targets, ot_inputs, compute_loss=True):
        sequence_output = self.uniter(input_ids, position_ids, img_feat, img_pos_feat, attention_mask,
                                      gather_index, output_all_encoded_layers=False)
        pooled_output = self.uniter.pooler(sequence_output)
        itm_scores = F_.SmallLinearFn.apply(pooled_output, self.itm_output.weight, self.itm_output.bias)

        # OT loss (model/pretrain.py:168-195): computed, then discarded exactly like the reference
        if ot_inputs is not None:
            ot_scatter = ot_inputs['ot_scatter']
            b = sequence_output.size(0)
            tl = input_ids.size(1)
            il = img_feat.size(1)
            max_l = max(ot_inputs['scatter_max'] + 1, tl + il)
            ot_scatter = ot_scatter.unsqueeze(-1).expand_as(sequence_output)
            ctx_emb = torch.zeros(b, max_l, self.config.hidden_size, dtype=sequence_output.dtype,
                                  device=sequence_output.device).scatter_(dim=1, index=ot_scatter,
                                                                          src=sequence_output)
            txt_emb = ctx_emb[:, :tl, :]
            img_emb = ctx_emb[:, tl:tl + il, :]
            txt_pad = ot_inputs['txt_pad']
            img_pad = ot_inputs['img_pad']
            ot_dist = optimal_transport_dist(txt_emb.float(), img_emb.float(), txt_pad, img_pad).to(txt_emb)
            ot_pos_dist = ot_dist.masked_select(targets == 1)
            ot_neg_dist = ot_dist.masked_select(targets == 0)
            ot_loss = (ot_pos_dist, ot_neg_dist)
        else:
            ot_loss = None
        self.last_ot_loss = ot_loss  # the reference drops it (returns are commented out, :199-203)

        if compute_loss:
            return F.cross_entropy(itm_scores, targets, reduction='none')
        return itm_scores

    def forward_mrc(self, input_ids, position_ids, img_feat, img_pos_feat, attention_mask, gather_index,
                    img_masks, img_mask_tgt, label_targets, task, compute_loss=True):
        sequence_output = self.uniter(input_ids, position_ids, img_feat, img_pos_feat, attention_mask,
                                      gather_index, output_all_encoded_layers=False, img_masks=img_masks)
        masked_output = self._compute_masked_hidden(sequence_output, img_mask_tgt)
        prediction_soft_label = self.region_classifier(masked_output)
        if compute_loss:
            if "kl" in task:
                prediction_soft_label = F.log_softmax(prediction_soft_label, dim=-1)
                return F.kl_div(prediction_soft_label, label_targets, reduction='none')
            # background class should not be the target
            label_targets = torch.max(label_targets[:, 1:], dim=-1)[1] + 1
            return F.cross_entropy(prediction_soft_label, label_targets, ignore_index=0, reduction='none')
        return prediction_soft_label
