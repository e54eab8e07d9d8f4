"""Evaluation / export path of the fine-tuning loop.

Mirrors what `TrainerTemplate.eval_model` / `export_*_predictions` and `data/metrics.py` do in the reference
(train_template.py:131-217, data/metrics.py:16-165): batched no-grad inference, sigmoid probabilities,
accuracy / precision / recall / F1 at a threshold, AUROC, the accuracy-optimal threshold and the
`id,proba,label[,gt]` CSV. What changes is the mechanics: the forward runs on the b200u kernels, probabilities
and the per-batch losses stay on the device until the end (the reference does `.cpu()` + `.item()` per batch,
train_template.py:121-126), AUROC is the rank statistic (what sklearn's roc_auc_score computes, ties get mid
ranks) and the threshold search is one sort + cumulative sums instead of one metrics pass per candidate.
"""
import torch

from . import functional as F_


@torch.no_grad()
def predict(model, batches, pos_wt=1.0):
    """Run `model` (MemeUniter, eval mode) over an iterable of device batch dicts.

    Returns (probs [N] f32, labels [N] or None, mean_loss float or None, ids [N] or None): probabilities are
    sigmoid(logits) (train_template.py:117-118), the loss is the mean of the per-batch BCE-with-logits losses
    like `eval_model` (train_template.py:145)."""
    was_training = model.training
    model.eval()
    probs, labels, losses, ids = [], [], [], []
    try:
        for b in batches:
            logits = model(input_ids=b["input_ids"], position_ids=b["position_ids"], img_feat=b["img_feat"],
                           img_pos_feat=b["img_pos_feat"], attention_mask=b["attn_mask"],
                           gather_index=b["gather_index"], output_all_encoded_layers=False)
            if b.get("labels") is not None:
                loss, _, p = F_.bce_with_logits(logits, b["labels"], pos_wt, want_grad=False)
                losses.append(loss)
                labels.append(b["labels"].reshape(-1))
            else:
                p = torch.sigmoid(logits.reshape(-1).float())
            probs.append(p)
            if b.get("ids") is not None:
                ids.append(b["ids"].reshape(-1))
    finally:
        model.train(was_training)
    probs = torch.cat(probs) if probs else torch.empty(0)
    labels = torch.cat(labels) if labels else None
    mean_loss = float(torch.stack(losses).mean().item()) if losses else None
    if probs.is_cuda:
        F_.check_input_errors()   # the host waits here anyway: surface out-of-range ids / gather indices
    return probs, labels, mean_loss, (torch.cat(ids) if ids else None)


def aucroc(probs, labels):
    """Area under the ROC curve = P(score_pos > score_neg) + 0.5 P(tie): the Mann-Whitney rank statistic with
    mid-ranks for ties (equals sklearn.metrics.roc_auc_score, data/metrics.py:151-165). 0.0 when only one class
    is present, like the reference."""
    probs = torch.as_tensor(probs).double().reshape(-1)
    labels = torch.as_tensor(labels).reshape(-1)
    pos = labels == 1
    n_pos, n_neg = int(pos.sum()), int((~pos).sum())
    if n_pos == 0 or n_neg == 0:
        return 0.0
    order = torch.argsort(probs)
    sp = probs[order]
    # mid-ranks: average 1-based rank over each run of equal scores
    n = sp.numel()
    idx = torch.arange(1, n + 1, dtype=torch.float64, device=sp.device)
    new_run = torch.ones(n, dtype=torch.bool, device=sp.device)
    new_run[1:] = sp[1:] != sp[:-1]
    run_id = torch.cumsum(new_run.long(), 0) - 1
    n_runs = int(run_id[-1]) + 1
    run_sum = torch.zeros(n_runs, dtype=torch.float64, device=sp.device).index_add_(0, run_id, idx)
    run_cnt = torch.zeros(n_runs, dtype=torch.float64, device=sp.device).index_add_(0, run_id, torch.ones_like(idx))
    ranks = (run_sum / run_cnt)[run_id]
    r_pos = ranks[pos[order]].sum()
    return float((r_pos - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


def _counts(probs, labels, threshold):
    preds = (probs > threshold).long()
    lab = labels.long()
    tp = ((preds == 1) & (lab == 1)).sum().float()
    tn = ((preds == 0) & (lab == 0)).sum().float()
    fp = ((preds == 1) & (lab == 0)).sum().float()
    fn = ((preds == 0) & (lab == 1)).sum().float()
    return tp, tn, fp, fn


def find_optimal_threshold(probs, labels):
    """Accuracy-optimal threshold with the reference's candidate list and tie-breaking
    (data/metrics.py:98-134): candidates 0.0, every probability in ascending order, 1.0; predictions are
    `probs > t`; the FIRST best candidate wins and, unless it is the first or last candidate, the threshold
    returned is the midpoint to the next candidate."""
    probs = torch.as_tensor(probs).reshape(-1)
    labels = torch.as_tensor(labels).reshape(-1).long()
    n = probs.numel()
    sp, order = torch.sort(probs)
    sl = labels[order]
    cand = torch.cat([probs.new_zeros(1), sp, probs.new_ones(1)])
    # for candidate t: predicted positive = #probs > t. With k = #probs <= t the correct count is
    # (#neg among the k smallest) + (#pos among the rest)
    k = torch.searchsorted(sp, cand, right=True)
    cneg = torch.cat([sl.new_zeros(1), torch.cumsum(1 - sl, 0)])
    cpos = torch.cat([sl.new_zeros(1), torch.cumsum(sl, 0)])
    correct = cneg[k] + (cpos[n] - cpos[k])
    acc = correct.double() / max(n, 1)
    best = int(torch.argmax(acc))          # first maximum
    # (torch.argmax returns the first maximal index for ties on CPU and CUDA)
    m = acc.max()
    best = int((acc == m).nonzero()[0])
    if best != cand.numel() - 1 and best != 0:
        return float((cand[best].double() + cand[best + 1].double()) / 2)
    return float(cand[best])


def standard_metrics_binary(probs, labels, threshold=0.5, add_aucroc=True, add_optimal_acc=False):
    """data/metrics.py:23-55: accuracy / recall / precision / F1 (+ aucroc, + optimal threshold & accuracy),
    plain floats. probs in [0, 1], labels in {0, 1}."""
    probs = torch.as_tensor(probs).reshape(-1)
    labels = torch.as_tensor(labels).reshape(-1)
    assert bool(torch.all((probs <= 1.0) & (probs >= 0.0))), "Probabilities must be between 0 and 1"
    assert bool(torch.all((labels == 0) | (labels == 1))), "Labels must be binary (0 or 1)"
    tp, tn, fp, fn = _counts(probs, labels, threshold)
    m = {}
    m["accuracy"] = float((tp + tn) / max(probs.numel(), 1))
    m["recall"] = float(tp / (tp + fn).clamp(min=1e-4))
    m["precision"] = float(tp / (tp + fp).clamp(min=1e-4))
    if m["recall"] == 0.0 or m["precision"] == 0.0:
        m["F1"] = 0.0
    else:
        m["F1"] = 2 * m["precision"] * m["recall"] / (m["precision"] + m["recall"])
    if add_aucroc:
        m["aucroc"] = aucroc(probs, labels)
    if add_optimal_acc:
        t = find_optimal_threshold(probs, labels)
        m["optimal_threshold"] = t
        m["optimal_accuracy"] = standard_metrics_binary(probs, labels, threshold=t, add_aucroc=False)["accuracy"]
    return m


def evaluate(model, batches, pos_wt=1.0):
    """`eval_model` (train_template.py:131-151): metrics dict (with the optimal-accuracy threshold) and the
    mean validation loss."""
    probs, labels, loss, _ = predict(model, batches, pos_wt)
    return standard_metrics_binary(probs, labels, add_optimal_acc=True), loss


def export_predictions(path, ids, probs, threshold=0.5, labels=None):
    """`_export_preds` (train_template.py:205-216): one line per sample, `id,proba,label[,gt]`."""
    ids = torch.as_tensor(ids).reshape(-1).cpu()
    probs = torch.as_tensor(probs).reshape(-1).float().cpu()
    preds = (probs > threshold).long()
    out = ["id,proba,label%s\n" % (",gt" if labels is not None else "")]
    lab = torch.as_tensor(labels).reshape(-1).cpu() if labels is not None else None
    for i in range(ids.shape[0]):
        line = "%i,%f,%i" % (int(ids[i]), float(probs[i]), int(preds[i]))
        if lab is not None:
            line += ",%i" % int(lab[i])
        out.append(line + "\n")
    with open(path, "w") as fh:
        fh.write("".join(out))
    return path
