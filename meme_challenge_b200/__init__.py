"""meme_challenge_b200 — B200-native (sm_100a) implementation of the UNITER forward/backward hot
path of Nithin-Holla/meme_challenge behind the reference's own module API.

    from meme_challenge_b200.model.model import UniterConfig, UniterModel
    from meme_challenge_b200.model.meme_uniter import MemeUniter
    from meme_challenge_b200.utils.utils import get_gather_index, get_attention_mask

Kernels live in csrc/ (hand-written CUDA: tcgen05+TMA GEMMs, fused attention, LayerNorm,
embeddings, gather, heads, IPOT, fused Adam) behind the C-ABI of include/b200u.h, bound with
ctypes in _lib.py. There is no CPU fallback.
"""
__version__ = "0.1.0"
