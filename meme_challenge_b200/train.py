"""Fused fine-tuning step for MemeUniter on b200u kernels, single GPU or data parallel.

Semantics are the reference trainer's (train_template.py:89-109, utils/optim_utils.py:16-46):
every micro-batch runs forward + `loss.backward()`; once `gradient_accumulation` micro-batches
are in, gradients are divided by the accumulation count, clipped to `max_grad_norm` (global L2),
Adam with L2 weight decay (no decay for names containing 'bias' / 'LayerNorm.bias' /
'LayerNorm.weight') updates the fp32 master weights, and the gradients are zeroed. What changes is
the mechanics: one flat gradient buffer, a three-launch optimizer (sum of squares, clip
coefficient, fused Adam that also refreshes the bf16 weight shadow), the whole step captured in
one CUDA graph, and — with world_size > 1 — one process per GPU with per-layer gradient buckets
all-reduced by NCCL over NVLink while the backward of earlier layers is still running (instead
of nn.DataParallel's per-step parameter broadcast + gradient reduce, train_template.py:58-59).
"""
import ctypes as C
import math
import os

import torch

from . import _lib, ops
from . import functional as F_
from .flat import FlatStore

P = _lib.ptr

NO_DECAY = ['bias', 'LayerNorm.bias', 'LayerNorm.weight']  # utils/optim_utils.py:16


def cosine_with_warmup(step, warmup_steps, total_steps):
    """transformers.get_cosine_schedule_with_warmup multiplier (train_template.py:80-82)."""
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    progress = float(step - warmup_steps) / float(max(1, total_steps - warmup_steps))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * progress)))


class GradBuckets(object):
    """Per-layer gradient buckets over ONE flat gradient buffer (no packing copies).

    `entries` is the flat layout [(parameter name, offset)] in module order. Bucket 0 covers the
    embeddings (everything before encoder layer 0), bucket i+1 covers encoder layer i, and the last
    bucket also takes whatever follows the last layer (pooler, classification head). Buckets are
    reduced with async all_reduce(SUM) as soon as the owning layer's backward is enqueued, i.e. in
    reverse layer order, and `wait()` joins them before the optimizer; the 1/world averaging is
    folded into the optimizer's gradient scale. Works on any backend (NCCL on GPUs, gloo in the
    CPU tests)."""

    def __init__(self, entries, flat_grad, process_group=None, world=1):
        self.grad = flat_grad
        self.pg = process_group
        self.world = world
        n = flat_grad.numel()
        first = {}
        for name, off in entries:
            if "encoder.layer." in name:
                l = int(name.split("encoder.layer.")[1].split(".")[0])
                first.setdefault(l, off)
        cuts = [first[l] for l in sorted(first)]
        bounds = [0] + cuts + [n]
        self.segments = [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)]
        self.pending = []
        self.reduced = []
        self.sync = False  # True: blocking collectives on the current stream (graph-capturable)

    def reduce_bucket(self, idx):
        lo, hi = self.segments[idx]
        self.reduced.append(idx)
        if hi <= lo or self.world <= 1:
            return
        if self.sync:
            torch.distributed.all_reduce(self.grad[lo:hi], group=self.pg)
        else:
            self.pending.append(torch.distributed.all_reduce(self.grad[lo:hi], group=self.pg, async_op=True))

    def wait(self):
        for w in self.pending:
            w.wait()
        self.pending = []
        order, self.reduced = self.reduced, []
        return order


class TrainStep(object):
    def __init__(self, model, lr=3e-5, weight_decay=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 gradient_accumulation=2, max_grad_norm=5.0, pos_wt=1.8, process_group=None,
                 overlap_comm=True, comm_sm_reserve=0, fuse_window=False):
        self.model = model
        self.um = model.uniter_model
        self.accum = int(gradient_accumulation)
        self.max_grad_norm = float(max_grad_norm)
        self.pos_wt = float(pos_wt)
        self.betas, self.eps = betas, eps
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.overlap_comm = overlap_comm
        # SMs left to NCCL while bucket all-reduces overlap the last micro-batch's backward: the
        # persistent kernels of that backward are sized for (SM count - reserve) so they stay one wave
        self.comm_sm_reserve = int(comm_sm_reserve)
        # software-pipeline the micro-batches of a window over two streams (see step()); B200U_PIPELINE=0 disables
        self.pipeline = os.environ.get("B200U_PIPELINE", "1") != "0"
        self._aux_stream = None
        # fuse_window: run the micro-batches of one accumulation window as ONE forward/backward pass over
        # accum * B samples (see _step_fused): same gradients, half the launches, twice the rows per GEMM
        self.fuse_window = bool(fuse_window)
        self._static_cat = None

        # one flat store for the whole MemeUniter (UNITER + classification head)
        store = FlatStore(model)
        self.um._store = store
        store.ensure()
        self.store = store
        dev = store.flat.device
        self.dev = dev
        n = store.flat.numel()
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.lr_t = torch.tensor([lr], device=dev, dtype=torch.float32)
        self.step_t = torch.zeros(1, device=dev, dtype=torch.int64)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.coef = torch.ones(1, device=dev, dtype=torch.float32)
        self.gnorm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.base_lr = lr
        self.host_step = 0

        # weight-decay runs over the flat layout
        starts, wds = [], []
        for name, p, off, cnt in store.entries:
            wd = 0.0 if any(nd in name for nd in NO_DECAY) else float(weight_decay)
            if not wds or wds[-1] != wd:
                starts.append(off)
                wds.append(wd)
        starts[0] = 0
        self.run_start = torch.tensor(starts + [n], device=dev, dtype=torch.int64)
        self.run_wd = torch.tensor(wds, device=dev, dtype=torch.float32)
        nchunks = (n + 1023) // 1024
        chunk_first = torch.arange(nchunks, dtype=torch.int64) * 1024
        cr = torch.searchsorted(torch.tensor(starts, dtype=torch.int64), chunk_first, right=True) - 1
        self.chunk_run = cr.to(torch.int32).to(dev)
        self.num_runs = len(wds)

        # gradient buckets: index 0 = embeddings, 1.. = encoder layers (the last also holds pooler + head)
        self.comm = GradBuckets([(e[0], e[2]) for e in store.entries], store.grad, process_group, self.world)
        self.buckets = self.comm.segments
        self.comm.sync = not overlap_comm
        # Word-embedding gradient [vocab, H] (20 % of all parameters, produced LAST by every backward, so its
        # dense all-reduce could not overlap anything, and row-sparse: <= B*T rows per micro-batch): with
        # world_size > 1 the micro-batches hand their touched rows over instead of scattering them, and after
        # the last backward the rows of the whole window are all-gathered and applied by every rank with a
        # deterministic segment add (sparse_word). No dense all-reduce of the table at all.
        self.word_slice = None
        for name, p, off, cnt in store.entries:
            if name.endswith("embeddings.word_embeddings.weight"):
                self.word_slice = (off, off + cnt, p)
        self.sparse_word = True
        self._word_rows = None
        self._word_dense = None
        self._graph = None
        self._static = None
        self.um._layer_grad_ready_cb = None
        store.refresh_shadow(force=True)

    # ------------------------------------------------------------------ buckets / comm
    def _allreduce_bucket(self, idx):
        self.comm.reduce_bucket(idx)

    def _on_layer_done(self, layer_idx):
        # called (from the autograd thread) once the backward of encoder layer `layer_idx` has been
        # enqueued: its bucket is final for this optimizer step, start the all-reduce now
        self.comm.reduce_bucket(layer_idx + 1)

    # ------------------------------------------------------------------ one micro-batch
    def _forward_loss(self, batch, last, first=True):
        """Forward + loss of one micro-batch on the current stream. Returns the state `_backward` needs.
        `first`: no earlier micro-batch of this window has accumulated gradients yet."""
        kw = dict(input_ids=batch["input_ids"], position_ids=batch["position_ids"],
                  img_feat=batch["img_feat"], img_pos_feat=batch["img_pos_feat"],
                  attention_mask=batch["attn_mask"], gather_index=batch["gather_index"],
                  output_all_encoded_layers=False)
        comm = self.world > 1 and last
        # the layer hooks are registered during the forward (they capture the callback), so it is only
        # set around this call
        self.um._layer_grad_ready_cb = self._on_layer_done if (comm and self.overlap_comm) else None
        # data parallel: EVERY micro-batch of the window hands its touched word-embedding rows over instead of
        # scattering them into the dense table; they are exchanged together after the last backward
        sparse = (self.world > 1 and self.overlap_comm and self.sparse_word and self.word_slice is not None)
        if sparse and first:
            self._word_rows = []
        self.um._sparse_word_cb = self._on_word_rows if sparse else None
        try:
            logits = self.model(**kw)
        finally:
            self.um._layer_grad_ready_cb = None
            self.um._sparse_word_cb = None
        loss, dlogits, probs = F_.bce_with_logits(logits, batch["labels"], self.pos_wt)
        return logits, dlogits, loss, probs, comm, sparse, first

    def _on_word_rows(self, d_rows, ids, pad):
        self._word_rows.append((d_rows, ids.contiguous(), pad))

    def _backward(self, state):
        logits, dlogits, loss, probs, comm, sparse, first = state
        dist = torch.distributed
        limit = comm and self.overlap_comm and self.comm_sm_reserve > 0
        if limit:
            sms = C.c_int()
            _lib.check(_lib.lib().b200u_device_info(C.byref(sms), None, None))
            _lib.lib().b200u_set_sm_limit(max(1, sms.value - self.comm_sm_reserve))
        try:
            torch.autograd.backward(logits, dlogits.view_as(logits))
        finally:
            if limit:
                _lib.lib().b200u_set_sm_limit(0)
        if comm:
            if not self.overlap_comm:
                for i in range(len(self.buckets) - 1, 0, -1):
                    self._allreduce_bucket(i)
            if sparse:
                self._finish_sparse_word()
            else:
                self._allreduce_bucket(0)

        return loss, probs

    def _finish_sparse_word(self):
        """Embedding bucket without the word table (dense, small) + the sparse word-row exchange."""
        dist = torch.distributed
        lo, hi, wp = self.word_slice
        b_lo, b_hi = self.buckets[0]
        self.comm.reduced.append(0)
        for a, b in ((b_lo, min(lo, b_hi)), (max(hi, b_lo), b_hi)):
            if b > a:
                self.comm.pending.append(dist.all_reduce(self.store.grad[a:b], group=self.pg, async_op=True))
        pad = self._word_rows[0][2]
        cur = torch.cuda.current_stream()
        for r, i, _ in self._word_rows:      # earlier micro-batches produced their rows on the other stream
            r.record_stream(cur)
            i.record_stream(cur)
        rows = torch.cat([r for r, _, _ in self._word_rows], 0)
        ids = torch.cat([i for _, i, _ in self._word_rows], 0)
        self._word_rows = None
        n, H = rows.shape
        all_rows = torch.empty(self.world * n, H, device=rows.device, dtype=rows.dtype)
        all_ids = torch.empty(self.world * n, device=ids.device, dtype=ids.dtype)
        dist.all_gather_into_tensor(all_rows, rows.contiguous(), group=self.pg)
        dist.all_gather_into_tensor(all_ids, ids, group=self.pg)
        # identical (all_rows, all_ids) on every rank + a deterministic, atomic-free segment add (rows sorted
        # by id, each run summed in order) => bit-identical word-embedding gradients on all replicas
        ids_sorted, perm = torch.sort(all_ids, stable=True)
        tot = self.world * n
        ops._call("b200u_embedding_segment_add", P(all_rows), P(ids_sorted), P(perm), P(self.store.grad[lo:hi]),
                  tot, H, C.c_longlong(pad))

    def micro_step(self, batch, last, first=True):
        return self._backward(self._forward_loss(batch, last, first))

    # ------------------------------------------------------------------ fused accumulation window
    @staticmethod
    def fuse_batches(batches):
        """Concatenate the micro-batches of one accumulation window along the batch dimension.

        Samples are independent in forward and backward (no cross-sample op on the path), so the window's
        gradient sum_i grad(mean-loss of micro-batch i) can be produced by one pass over all samples.
        Micro-batches padded to different widths are right-padded to the widest: attention-mask columns
        with 0 (never attended), gather-index columns with the identity tail the reference's
        get_gather_index produces for padded positions (utils/utils.py:113), region rows with zeros."""
        L = max(b["attn_mask"].shape[1] for b in batches)
        R = max(b["img_feat"].shape[1] for b in batches)
        out = {}
        for k in ("input_ids", "position_ids", "img_feat", "img_pos_feat", "attn_mask", "gather_index", "labels"):
            parts = []
            for b in batches:
                t = b[k]
                if k in ("img_feat", "img_pos_feat") and t.shape[1] < R:
                    t = torch.cat([t, t.new_zeros(t.shape[0], R - t.shape[1], t.shape[2])], 1)
                elif k == "attn_mask" and t.shape[1] < L:
                    t = torch.cat([t, t.new_zeros(t.shape[0], L - t.shape[1])], 1)
                elif k == "gather_index" and t.shape[1] < L:
                    tail = torch.arange(t.shape[1], L, device=t.device, dtype=t.dtype).unsqueeze(0)
                    t = torch.cat([t, tail.expand(t.shape[0], -1)], 1)
                parts.append(t)
            out[k] = torch.cat(parts, 0)
        return out

    def _step_fused(self, batches, cat=None):
        """All micro-batches of the window in one pass. Per-micro-batch mean losses (and their gradients)
        are formed on the slices of the joint logits, so the accumulated gradient, the 1/accum averaging
        and the reported losses are those of the sequential window (train_template.py:95-103)."""
        if cat is None:
            cat = self.fuse_batches(batches)
        sizes = [b["labels"].shape[0] for b in batches]
        kw = dict(input_ids=cat["input_ids"], position_ids=cat["position_ids"], img_feat=cat["img_feat"],
                  img_pos_feat=cat["img_pos_feat"], attention_mask=cat["attn_mask"],
                  gather_index=cat["gather_index"], output_all_encoded_layers=False)
        comm = self.world > 1
        self.um._layer_grad_ready_cb = self._on_layer_done if (comm and self.overlap_comm) else None
        sparse = (comm and self.overlap_comm and self.sparse_word and self.word_slice is not None)
        if sparse:
            self._word_rows = []
        self.um._sparse_word_cb = self._on_word_rows if sparse else None
        try:
            logits = self.model(**kw)
        finally:
            self.um._layer_grad_ready_cb = None
            self.um._sparse_word_cb = None
        flat_logits = logits.reshape(-1)
        dl = torch.empty_like(flat_logits, dtype=torch.float32)
        outs, o = [], 0
        for n in sizes:
            loss, _, probs = F_.bce_with_logits(flat_logits[o:o + n], cat["labels"][o:o + n], self.pos_wt,
                                                dl_out=dl[o:o + n])
            outs.append((loss, probs))
            o += n
        state = (logits, dl, None, None, comm, sparse, True)
        self._backward(state)
        return outs

    # ------------------------------------------------------------------ optimizer
    def optimizer_step(self):
        self.comm.wait()
        g = self.store.grad
        n = g.numel()
        self.sumsq.zero_()
        ops._call("b200u_grad_sumsq", P(g), C.c_size_t(n), P(self.sumsq))
        pre = 1.0 / (self.accum * self.world)  # average_gradients + mean over ranks
        ops._call("b200u_clip_coef", P(self.sumsq), pre, self.max_grad_norm, P(self.coef), P(self.gnorm))
        ops.counter_add(self.step_t, 1)
        ops._call("b200u_adam_step", P(self.store.flat), P(g), P(self.m), P(self.v), P(self.store.shadow),
                  C.c_size_t(n), P(self.run_start), P(self.run_wd), P(self.chunk_run), self.num_runs,
                  P(self.coef), P(self.lr_t), P(self.step_t), self.betas[0], self.betas[1], self.eps, 1)

    def set_lr(self, lr):
        self.lr_t.fill_(lr)

    # ------------------------------------------------------------------ public: one optimizer step
    def step(self, batches):
        """batches: list of `gradient_accumulation` device batch dicts. Returns the list of
        (loss[1], probs[B]) device tensors of the micro-batches.

        The micro-batches of one accumulation window are independent until their gradients meet in the
        flat buffer, so the window is software-pipelined over two streams: the forward of micro-batch
        i+1 runs beside the backward of micro-batch i (forward chains and backward chains each leave
        SMs idle in their launch / first-load / epilogue-drain phases; two chains fill each other's
        gaps). Backward passes stay ordered (they share scratch buffers and, with world_size > 1, the
        last one triggers the bucket all-reduces), forwards stay ordered (dropout seed sequence)."""
        # (the reference's first optimizer step fires after ONE micro-batch and still divides by the
        #  accumulation count, train_template.py:101-103: a shorter window is allowed, the scale is not changed)
        assert 1 <= len(batches) <= self.accum
        n = len(batches)
        if self.fuse_window and n > 1:
            outs = self._step_fused(batches, self._static_cat if batches is self._static else None)
        elif n == 1 or not self.pipeline:
            outs = [self.micro_step(b, last=(i == n - 1), first=(i == 0)) for i, b in enumerate(batches)]
        else:
            main = torch.cuda.current_stream()
            if self._aux_stream is None:
                self._aux_stream = torch.cuda.Stream()
            streams = [main, self._aux_stream]
            state, outs = {}, [None] * n

            def fwd(i, after):
                s = streams[i % 2]
                if after is not None:
                    s.wait_event(after)
                with torch.cuda.stream(s):
                    state[i] = self._forward_loss(batches[i], last=(i == n - 1), first=(i == 0))
                    return s.record_event()

            def bwd(i, after):
                s = streams[i % 2]
                if after is not None:
                    s.wait_event(after)
                with torch.cuda.stream(s):
                    outs[i] = self._backward(state.pop(i))
                    return s.record_event()

            ev_f = fwd(0, None)
            ev_b = None
            for i in range(n):
                ev_next = fwd(i + 1, ev_f) if i + 1 < n else None
                ev_b = bwd(i, ev_b)
                ev_f = ev_next
            main.wait_event(ev_b)     # join before the optimizer (and before the caller reads the outputs)
        self.optimizer_step()
        self.host_step += 1
        return outs

    # ------------------------------------------------------------------ CUDA-graph replay
    def capture(self, example_batches, warmup=3):
        """Capture one full optimizer step (all micro-batches + optimizer) in a CUDA graph over
        static input buffers. Returns the static buffers; fill them and call replay().

        With world_size > 1 the bucket all-reduces are captured too: the collectives issued from the
        backward hooks fork onto the process group's stream inside the capture and `wait()` joins
        them before the optimizer nodes, so the replayed graph keeps the comm/backward overlap.
        (thread_local capture mode: the NCCL watchdog thread may poll events while we capture.)"""
        if self.fuse_window and len(example_batches) > 1:
            # the static micro-batch buffers are views of one concatenated set: refreshing them fills the
            # fused batch in place (micro-batches of one captured window share their padded widths)
            ex = [{k: v for k, v in b.items() if torch.is_tensor(v)} for b in example_batches]
            cat = self.fuse_batches(ex)
            static, o = [], 0
            for b in ex:
                n = b["labels"].shape[0]
                static.append({k: v[o:o + n] for k, v in cat.items()})
                o += n
            self._static_cat = cat
        else:
            static = [{k: v.clone() for k, v in b.items() if torch.is_tensor(v)} for b in example_batches]
            self._static_cat = None
        self._static = static
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.step(static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        mode = "thread_local" if self.world > 1 else "global"
        with torch.cuda.graph(graph, capture_error_mode=mode):
            outs = self.step(static)
        self._graph, self._static, self._static_out = graph, static, outs
        return static

    def load_static(self, batches):
        for dst, src in zip(self._static, batches):
            for k, v in dst.items():
                v.copy_(src[k], non_blocking=True)

    def replay(self):
        self._graph.replay()
        self.host_step += 1
        return self._static_out
