"""Flat parameter / gradient / bf16-shadow storage for a module tree.

The reference keeps 212 separate fp32 tensors and lets torch touch each one in the optimizer,
in clip_grad_norm_ and (under nn.DataParallel) in the per-step broadcast/reduce
(train_template.py:58-59,89-107). Here every parameter of the wrapped module is a *view* into one
fp32 buffer, its `.grad` a view into a second one, and the tcgen05 GEMMs read a third, bf16
"shadow" copy. That gives
  * a fused [3H, H] query/key/value weight without changing the state_dict (the three
    nn.Linear weights are adjacent views),
  * wgrad kernels that accumulate straight into `.grad` (no per-tensor copies),
  * gradient buckets for NCCL that are plain slices of one buffer (no packing),
  * a one-launch fused Adam.
The state_dict layout (names, shapes, fp32 dtype) is untouched: model/model.py:148-214 and
utils/save.py:53-64 keep working.
"""
import torch

from . import _lib, ops

import weakref

ALIGN = 8  # elements: keeps every tensor 32-byte (fp32) / 16-byte (bf16) aligned

_STORE_OF = {}  # id(param) -> weakref(FlatStore) for parameters that live in a flat store


def store_of(p):
    ref = _STORE_OF.get(id(p))
    st = ref() if ref is not None else None
    if st is not None and st.flat is not None and id(p) in st.index and st.valid():
        return st
    return None


def _ordered_params(root):
    """named_parameters() order, except each attention block lists q/k/v weights (then biases)
    back to back so they form one [3H,H] (resp. [3H]) slab."""
    from .model.layer import BertSelfAttention

    fused = {}
    for mod_name, mod in root.named_modules():
        if isinstance(mod, BertSelfAttention):
            pre = mod_name + "." if mod_name else ""
            names = [pre + n for n in ("query.weight", "key.weight", "value.weight",
                                       "query.bias", "key.bias", "value.bias")]
            fused[names[0]] = names
            for n in names[1:]:
                fused[n] = None
    named = dict(root.named_parameters())
    out, seen = [], set()
    for name, p in named.items():
        if id(p) in seen:
            continue
        grp = fused.get(name, [name])
        if grp is None:
            continue
        for n in grp:
            q = named[n]
            if id(q) not in seen:
                seen.add(id(q))
                out.append((n, q))
    return out


class FlatStore(object):
    def __init__(self, root):
        self.root = root
        self.flat = None
        self.grad = None
        self.shadow = None
        self.entries = []        # (name, param, offset, numel)
        self.index = {}          # id(param) -> (offset, numel)
        self._shadow_version = None
        self._ptr_sig = None
        self.touched = set()     # id(param) of every parameter a backward kernel has accumulated a gradient into

    # ------------------------------------------------------------------ build / validate
    def _signature(self):
        return tuple(p.data_ptr() for _, p in self.root.named_parameters())

    def valid(self):
        return self.flat is not None and self._ptr_sig == self._signature()

    def ensure(self):
        if not self.valid():
            self.build()
        return self

    def build(self):
        params = _ordered_params(self.root)
        if not params:
            raise _lib.B200UError("FlatStore: module has no parameters")
        dev = params[0][1].device
        if dev.type != "cuda":
            raise _lib.B200UError(
                "b200u modules run on CUDA (sm_100a) only — move the model with .to('cuda') first; "
                "there is no CPU fallback")
        total, offs = 0, []
        for _, p in params:
            if p.dtype != torch.float32:
                raise _lib.B200UError("FlatStore expects fp32 master parameters (got %s)" % p.dtype)
            offs.append(total)
            total += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.entries, self.index = [], {}
        with torch.no_grad():
            for (name, p), off in zip(params, offs):
                n = p.numel()
                view = flat[off:off + n].view(p.shape)
                view.copy_(p.data)
                gview = grad[off:off + n].view(p.shape)
                if p.grad is not None:
                    gview.copy_(p.grad)
                p.data = view
                p.grad = gview
                self.entries.append((name, p, off, n))
                self.index[id(p)] = (off, n)
                _STORE_OF[id(p)] = weakref.ref(self)
        self.flat, self.grad = flat, grad
        self.shadow = torch.empty(total, device=dev, dtype=torch.bfloat16)
        self._shadow_version = None
        self._ptr_sig = self._signature()
        return self

    # ------------------------------------------------------------------ views
    def _view(self, buf, p, shape=None):
        off, n = self.index[id(p)]
        return buf[off:off + n].view(p.shape if shape is None else shape)

    def w16(self, p):
        """bf16 shadow of parameter p (refreshed lazily)."""
        return self._view(self.shadow, p)

    def g32(self, p):
        return self._view(self.grad, p)

    def fused(self, buf, first, rows):
        """[rows, ...] slab starting at parameter `first` (q/k/v fusion)."""
        off, n = self.index[id(first)]
        per_row = n // first.shape[0]
        return buf[off:off + rows * per_row].view((rows,) + tuple(first.shape[1:]))

    # ------------------------------------------------------------------ shadow / grads
    def version_sig(self):
        # in-place updates through optimizers / load_state_dict / copy_ bump each param's version
        return sum(e[1]._version for e in self.entries)

    def refresh_shadow(self, force=False):
        v = self.version_sig()
        if force or self._shadow_version != v:
            ops.cast_f32_to_bf16(self.flat, self.shadow)
            self._shadow_version = v

    def mark_shadow_current(self):
        self._shadow_version = self.version_sig()

    def attach_grads(self):
        """Make every param.grad the flat-buffer view again (after zero_grad(set_to_none=True))."""
        missing = [e for e in self.entries if e[1].grad is None or
                   e[1].grad.data_ptr() != self.grad.data_ptr() + 4 * e[2]]
        if not missing:
            return
        if len(missing) == len(self.entries) and all(e[1].grad is None for e in missing):
            self.grad.zero_()
        for name, p, off, n in missing:
            gv = self.grad[off:off + n].view(p.shape)
            if p.grad is None:
                if len(missing) != len(self.entries):
                    gv.zero_()
            else:
                gv.copy_(p.grad)
            p.grad = gv

    def zero_grad(self):
        self.grad.zero_()
