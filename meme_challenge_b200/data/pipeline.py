"""Input pipeline of the fine-tuning hot path.

Mirrors what the reference does on the host for every batch — region-feature loading and 7-d box
construction (data/dataset_template.py:92-114), the MemeDataset collate (data/meme_dataset.py:152-214:
zero-padded features, position ids, attention mask, gather index) — and replaces its blocking, pageable
`.to(device)` per tensor (train_template.py:397-399) with a pinned, double-buffered prefetcher whose
host->device copies run on their own stream while the previous step computes.
"""
import os

import numpy as np
import torch

from ..utils.utils import get_attention_mask, get_gather_index

BATCH_KEYS = ("input_ids", "position_ids", "img_feat", "img_pos_feat", "attn_mask", "gather_index", "labels")


def box7(bbox, img_width=None, img_height=None, normalize=False):
    """[n, 4] (x1, y1, x2, y2) -> [n, 7] (x1, y1, x2, y2, w, h, w*h), the UNITER location feature
    (data/dataset_template.py:100-113). numpy or torch input; `normalize` divides by the image size first."""
    t = torch.as_tensor(bbox).clone()
    x1, y1, x2, y2 = t[:, 0:1], t[:, 1:2], t[:, 2:3], t[:, 3:4]
    if normalize:
        x1 = x1 / img_width
        x2 = x2 / img_width
        y1 = y1 / img_height
        y2 = y2 / img_height
    w = x2 - x1
    h = y2 - y1
    return torch.cat((x1, y1, x2, y2, w, h, w * h), dim=1)


def load_img_feature(feature_dir, img_id, normalize=False):
    """data/dataset_template.py:92-114: `<id>.npy` holds the [num_bb, 2048] region features, `<id>_info.npy`
    a pickled dict with `bbox`, `image_width`, `image_height`, `objects` and `objects_conf` / `cls_prob`.
    Returns (img_feat, img_pos_feat, objects, objects_conf) like the reference."""
    img_id = str(img_id).zfill(5)
    img_feat = torch.from_numpy(np.load(os.path.join(feature_dir, "%s.npy" % img_id)))
    info = np.load(os.path.join(feature_dir, "%s_info.npy" % img_id), allow_pickle=True).item()
    conf = info["objects_conf"] if "objects_conf" in info else info["cls_prob"].max(axis=-1)
    pos = box7(info["bbox"], info["image_width"], info["image_height"], normalize)
    return img_feat, pos, info["objects"], conf


def collate_memes(samples, input_ids, text_len, pad_regions_are_valid=True, device=None, token_type_ids=None):
    """The compact-batch branch of MemeDataset.get_collate_fn (data/meme_dataset.py:152-214).

    samples: list of dicts with `img_feat` [nbb_i, D], `img_pos_feat` [nbb_i, 7], `label`, optionally
    `data_id`; input_ids: [B, T] tokenised text already padded to max_length (train_uniter.py:125);
    text_len: true token counts. pad_regions_are_valid=True reproduces the reference exactly: it takes
    `img_len` from the ALREADY padded feature tensor (meme_dataset.py:186), so every sample's mask covers
    the batch-maximum number of regions; False masks each sample's own region count (SURVEY.md §8d's
    ragged synthetic batches). Mask and gather index are built on `device` when given."""
    feats = [torch.as_tensor(s["img_feat"]) for s in samples]
    poss = [torch.as_tensor(s["img_pos_feat"]) for s in samples]
    img_feat = torch.nn.utils.rnn.pad_sequence(feats, batch_first=True, padding_value=0)
    img_pos_feat = torch.nn.utils.rnn.pad_sequence(poss, batch_first=True, padding_value=0)
    labels = torch.stack([torch.as_tensor(s["label"]) for s in samples], dim=0)
    text_len = [int(t) for t in text_len]
    B, T = input_ids.shape
    position_ids = torch.arange(0, T, device=input_ids.device).unsqueeze(0).repeat(B, 1)
    img_len = [img_feat.shape[1]] * B if pad_regions_are_valid else [f.shape[0] for f in feats]
    attn_mask = get_attention_mask(text_len, img_len, device=device)
    gather_index = get_gather_index(text_len, img_len, B, T, attn_mask.shape[1], device=device)
    batch = {"input_ids": input_ids, "position_ids": position_ids, "img_feat": img_feat,
             "img_pos_feat": img_pos_feat, "token_type_ids": token_type_ids, "attn_mask": attn_mask,
             "gather_index": gather_index, "labels": labels}
    if all("data_id" in s for s in samples):
        batch["ids"] = torch.stack([torch.as_tensor(s["data_id"]) for s in samples], dim=0)
    return batch


class PinnedPrefetcher(object):
    """Double-buffered host -> device feeder.

    `source` yields host items: a batch dict of CPU tensors, or a list of them (the micro-batches of one
    accumulation window). Each of the `depth` slots owns pinned host staging buffers and device buffers;
    the copies of item i+1 are issued on a private copy stream while the consumer works on item i, and the
    consumer's stream waits on the slot's event — never on the host. A slot is overwritten only after the
    work the consumer enqueued for it has finished (event recorded on the consumer's stream when the next
    item is requested). Non-tensor entries and `None`s pass through. labels are converted to float32 (the
    dtype the BCE kernel reads, train_template.py:99 `labels.float()`)."""

    def __init__(self, source, device, depth=2, auto_prefetch=True):
        if not torch.cuda.is_available():
            from .._lib import B200UError
            raise B200UError("PinnedPrefetcher needs a CUDA device (no CPU fallback)")
        self.source = iter(source)
        self.device = torch.device(device)
        self.depth = max(2, int(depth))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [dict(pinned={}, dev={}, ready=None, consumed=None) for _ in range(self.depth)]
        self.queue = []          # slot indices with copies in flight, oldest first
        self.next_slot = 0
        self.h2d_bytes = 0       # bytes copied for the most recent item
        self._exhausted = False
        # auto_prefetch=False: the consumer calls prefetch_next() itself, AFTER it has launched the step that
        # consumes the current item, so the host time of issuing the next copies hides behind that step
        self.auto_prefetch = bool(auto_prefetch)
        self._pinned_ok = {}     # data_ptr -> bool (Tensor.is_pinned() queries the driver: cache per source buffer)

    # ------------------------------------------------------------------ internals
    def _buf(self, slot, name, t):
        key = (name, tuple(t.shape), t.dtype)
        hit = slot["pinned"].get(name)
        if hit is None or hit[0] != key:
            pinned = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            dev = torch.empty(t.shape, dtype=t.dtype, device=self.device)
            hit = (key, pinned, dev)
            slot["pinned"][name] = hit
        return hit[1], hit[2]

    def _stage(self, slot, prefix, batch):
        out, nbytes = {}, 0
        for k, v in batch.items():
            if not torch.is_tensor(v):
                out[k] = v
                continue
            if k == "labels" and v.dtype != torch.float32:
                v = v.float()
            pinned, dev = self._buf(slot, prefix + k, v)
            key = (v.data_ptr(), v.numel())
            is_pinned = self._pinned_ok.get(key)
            if is_pinned is None:
                is_pinned = self._pinned_ok[key] = bool(v.is_pinned())
            if is_pinned:
                src = v
            else:
                pinned.copy_(v)
                src = pinned
            dev.copy_(src, non_blocking=True)
            nbytes += v.numel() * v.element_size()
            out[k] = dev
        return out, nbytes

    def _issue(self):
        if self._exhausted:
            return False
        try:
            item = next(self.source)
        except StopIteration:
            self._exhausted = True
            return False
        idx = self.next_slot
        self.next_slot = (idx + 1) % self.depth
        slot = self.slots[idx]
        if slot["consumed"] is not None:
            self.copy_stream.wait_event(slot["consumed"])   # the work that read this slot has finished
        with torch.cuda.stream(self.copy_stream):
            if isinstance(item, (list, tuple)):
                staged, total = [], 0
                for i, b in enumerate(item):
                    o, n = self._stage(slot, "%d." % i, b)
                    staged.append(o)
                    total += n
            else:
                staged, total = self._stage(slot, "", item)
            slot["ready"] = self.copy_stream.record_event()
        slot["item"], slot["bytes"] = staged, total
        self.queue.append(idx)
        return True

    # ------------------------------------------------------------------ iteration
    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        # everything the consumer enqueued so far may still read the slots handed out earlier
        done = cur.record_event()
        for s in self.slots:
            if s.get("handed"):
                s["consumed"] = done
                s["handed"] = False
        while len(self.queue) < self.depth - 1 or not self.queue:
            if not self._issue():
                break
        if not self.queue:
            raise StopIteration
        idx = self.queue.pop(0)
        slot = self.slots[idx]
        cur.wait_event(slot["ready"])
        slot["handed"] = True
        self.h2d_bytes = slot["bytes"]
        if self.auto_prefetch:
            self._issue()   # keep the pipeline full: the next item's copies start now
        return slot["item"]

    def prefetch_next(self):
        """Issue the copies of the next item now (auto_prefetch=False: call it right after launching the work
        that consumes the current item)."""
        if len(self.queue) < self.depth - 1:
            self._issue()
