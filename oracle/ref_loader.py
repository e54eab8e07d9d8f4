"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference modules from oracle/_ref/ (built by
oracle/build_ref.py) or, in the build container, straight from /root/reference.

Shims, applied to the running interpreter and never to the files:
  * `apex.normalization.fused_layer_norm.FusedLayerNorm` -> `torch.nn.LayerNorm` (apex is not installed;
    same arithmetic: biased variance, eps inside the sqrt, affine). Installed BEFORE the reference is
    imported so `init_weights`' isinstance check (model/model.py:142) sees the same class.
  * `model.ot.trace` -> diagonal sum: the reference builds a uint8 mask and `masked_select` raises on
    torch >= 2 (model/ot.py:24-32); same elements, same order.
Only tests/, bench.py's reference / cpu_baseline legs and __graft_entry__ may import this module.
"""
import importlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}


def reference_root():
    """Directory that holds the reference's model/ package: oracle/_ref if built, else the checkout."""
    ref = os.path.join(HERE, "_ref")
    if os.path.exists(os.path.join(ref, "model", "model.py")):
        return ref
    chk = os.environ.get("B200U_REFERENCE", "/root/reference")
    if os.path.exists(os.path.join(chk, "model", "model.py")):
        return chk
    return None


def available():
    return reference_root() is not None


def load():
    """Returns a namespace with the reference modules: .model, .layer, .meme_uniter, .ot, .pretrain,
    .optim_utils and .root. Raises RuntimeError when neither oracle/_ref nor the checkout exists."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference modules unavailable: run `python oracle/build_ref.py` where /root/reference exists")
    apex = types.ModuleType("apex")
    norm = types.ModuleType("apex.normalization")
    fln = types.ModuleType("apex.normalization.fused_layer_norm")
    fln.FusedLayerNorm = torch.nn.LayerNorm
    sys.modules.setdefault("apex", apex)
    sys.modules.setdefault("apex.normalization", norm)
    sys.modules.setdefault("apex.normalization.fused_layer_norm", fln)
    # the reference's top-level package names are `model` and `utils`: import them under those names
    # from `root`, then restore sys.path / sys.modules so nothing else resolves against the reference
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ("model", "utils")}
    for k in list(sys.modules):
        if k == "model" or k.startswith("model.") or k == "utils" or k.startswith("utils."):
            del sys.modules[k]
    sys.path.insert(0, root)
    try:
        ns = types.SimpleNamespace(root=root)
        ns.model = importlib.import_module("model.model")
        ns.layer = importlib.import_module("model.layer")
        ns.meme_uniter = importlib.import_module("model.meme_uniter")
        ns.ot = importlib.import_module("model.ot")
        ns.pretrain = importlib.import_module("model.pretrain")
        ns.optim_utils = importlib.import_module("utils.optim_utils")
        ns.ot.trace = lambda x: torch.diagonal(x, dim1=-2, dim2=-1).sum(-1)
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k == "model" or k.startswith("model.") or k == "utils" or k.startswith("utils."):
                del sys.modules[k]
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v
    _CACHE["ns"] = ns
    return ns
