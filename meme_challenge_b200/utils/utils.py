"""Index / mask construction of the hot path (reference utils/utils.py:111-141), bit-exact.

The three functions keep the reference names, argument meaning and results (int64 gather index,
float32 0/1 mask, zero-padded feature batches); they are vectorised instead of looping in Python
and accept a `device` so the collate can build them where the batch lives.
"""
import torch


def get_gather_index(txt_lens, num_bbs, batch_size, max_len, out_size, device=None):
    """utils/utils.py:111-117: idx[i, j] = j, except idx[i, tl_i : tl_i+nbb_i] = max_len + (0..nbb_i-1)."""
    assert len(txt_lens) == len(num_bbs) == batch_size
    tl = torch.as_tensor(list(txt_lens), dtype=torch.long, device=device).unsqueeze(1)
    nbb = torch.as_tensor(list(num_bbs), dtype=torch.long, device=device).unsqueeze(1)
    j = torch.arange(0, out_size, dtype=torch.long, device=device).unsqueeze(0).repeat(batch_size, 1)
    img = (j >= tl) & (j < tl + nbb)
    return torch.where(img, j - tl + max_len, j)


def get_attention_mask(text_len, img_len, device=None):
    """utils/utils.py:120-125: ones(tl_i + nbb_i) right-padded with 0 to the batch maximum (float32)."""
    tot = torch.as_tensor([int(t) + int(i) for t, i in zip(text_len, img_len)], dtype=torch.long,
                          device=device)
    width = int(tot.max().item()) if tot.numel() else 0
    j = torch.arange(width, dtype=torch.long, device=device).unsqueeze(0)
    return (j < tot.unsqueeze(1)).to(torch.float32)


def pad_tensors(tensors, lens=None, pad=0):
    """utils/utils.py:128-141: B x [T, ...] -> zero (or `pad`) padded [B, max_len, hid]."""
    if lens is None:
        lens = [t.size(0) for t in tensors]
    max_len = max(lens)
    bs = len(tensors)
    hid = tensors[0].size(-1)
    dtype = tensors[0].dtype
    output = torch.zeros(bs, max_len, hid, dtype=dtype)
    if pad:
        output.data.fill_(pad)
    for i, (t, l) in enumerate(zip(tensors, lens)):
        output.data[i, :l, ...] = t.data
    return output
