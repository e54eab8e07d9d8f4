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
import sys

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

    def __init__(self, entries, flat_grad, process_group=None, world=1, comm_dtype=None, comm_impl="auto",
                 tail_lo=None):
        self.grad = flat_grad
        self.pg = process_group
        self.world = world
        # comm_dtype=torch.bfloat16: the encoder-layer buckets (everything from layer 0 on) are cast to a
        # bf16 staging buffer when their layer's backward is done and all-reduced THERE: half the NVLink
        # volume; the fp32 accumulation over the window stays local and the optimizer reads the reduced
        # bf16 values (b200u_adam_step g16 range). The embedding bucket stays fp32.
        self.comm_dtype = comm_dtype
        self.g16 = None
        self.g16_lo = 0
        self._cast_stream = None
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
        self._side_work = False
        self.sync = False  # True: blocking collectives on the current stream (graph-capturable)
        self.peer = None       # copy-engine exchange state (see _setup_peer)
        if comm_dtype == torch.bfloat16 and world > 1 and len(self.segments) > 1:
            self.g16_lo = self.segments[1][0]
            want_ce = comm_impl in ("auto", "ce") and flat_grad.is_cuda and \
                os.environ.get("B200U_DP_IMPL", "ce") != "nccl"
            if want_ce:
                # tail_lo (end of the word-embedding table, whose rows travel sparsely): the dense rest of the
                # embedding bucket [tail_lo, first layer) joins the bf16 symmetric range as one more bucket
                tail = tail_lo if (tail_lo is not None and self.segments[0][0] <= tail_lo < self.segments[0][1]
                                   and tail_lo % 8 == 0) else None
                err = None
                try:
                    self._setup_peer(n, tail)
                except Exception as e:   # no P2P / symmetric memory on this box: NCCL path
                    err = e
                # every rank must take the same path: agree on the outcome
                ok = torch.tensor([0 if err is not None else 1], device=flat_grad.device, dtype=torch.int32)
                torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN, group=process_group)
                if int(ok.item()) == 0:
                    if comm_impl == "ce":
                        raise err if err is not None else RuntimeError("symmetric memory unavailable on a peer rank")
                    if torch.distributed.get_rank(process_group) == 0:
                        msg = (str(err).splitlines()[0] if err is not None and str(err) else type(err).__name__)
                        print("b200u: symmetric-memory gradient exchange unavailable (%s); using NCCL all-reduce" % msg,
                              file=sys.stderr)
                    self.peer = None
                    self.g16 = None
                    self.g16_lo = self.segments[1][0]
            if self.peer is None:
                self.g16 = torch.zeros(n - self.g16_lo, device=flat_grad.device, dtype=torch.bfloat16)

    # ------------------------------------------------------------------ copy-engine all-reduce over NVLink
    @staticmethod
    def slice_plan(segments, world, n_total, tail_lo=None):
        """Layout of the copy-engine exchange: (first element of the bf16 range, its length, per-bucket
        (offset in the bf16 range, bucket length, slice length, staging offset) for the encoder-layer buckets [+ the
        dense tail of the embedding bucket as the LAST entry], staging elements). Rank r owns elements
        [r * slice, (r + 1) * slice) of a bucket (clipped to its length); slice lengths are multiples of 8 elements
        so every copy starts 16-byte aligned; a bucket's staging area holds one slice per source rank."""
        g16_lo = segments[1][0] if tail_lo is None else tail_lo
        ranges = list(segments[1:]) + ([(tail_lo, segments[0][1])] if tail_lo is not None else [])
        plan, stage_off = [], 0
        for (lo, hi) in ranges:
            ln = hi - lo
            sl = ((ln + world - 1) // world + 7) // 8 * 8
            plan.append((lo - g16_lo, ln, sl, stage_off))
            stage_off += world * sl
        return g16_lo, n_total - g16_lo, plan, stage_off

    def _setup_peer(self, n_total, tail_lo=None):
        """Two-shot all-reduce of the bf16 layer buckets WITHOUT collective kernels on the SMs.

        An NCCL ring all-reduce keeps 8-32 CTAs resident for ~80 us per layer bucket while the backward's
        persistent tcgen05 GEMMs want all 148 SMs: measured on 2 B200s it costs ~20 us of GEMM time per layer
        (260 us per step, the whole data-parallel loss) whatever the channel count. Here the bucket lives in
        symmetric memory (every rank maps every peer's buffer, NVSwitch gives all pairs full bandwidth) and
        moves with the COPY ENGINES: rank r owns slice r of each bucket;
          1. every rank pushes slice p of its bucket into peer p's staging area        (N-1 cudaMemcpyAsync P2P)
          2. barrier; rank r sums its own slice with the N-1 staged copies             (b200u_slice_sum_bf16)
          3. rank r pushes the reduced slice r into every peer's bucket                (N-1 cudaMemcpyAsync P2P)
          4. barrier.
        All of it is stream-ordered on the side stream that already does the fp32 -> bf16 cast, CUDA-graph
        capturable (memcpy + kernel nodes), and deterministic: slice r is summed by one rank in a fixed order
        and broadcast, so replicas stay bit-identical."""
        import torch.distributed._symmetric_memory as symm_mem
        dist = torch.distributed
        W, dev = self.world, self.grad.device
        rank = dist.get_rank(self.pg)
        assert W - 1 <= 15, "b200u_slice_sum_bf16 takes at most 15 peers"
        group = self.pg if self.pg is not None else dist.group.WORLD
        g16_lo, n16, plan, stage_off = self.slice_plan(self.segments, W, n_total, tail_lo)
        g16 = symm_mem.empty(n16, dtype=torch.bfloat16, device=dev)
        stage = symm_mem.empty(max(stage_off, 8), dtype=torch.bfloat16, device=dev)
        g16.zero_()
        h_g = symm_mem.rendezvous(g16, group)
        h_s = symm_mem.rendezvous(stage, group)
        peers_g = [g16 if p == rank else h_g.get_buffer(p, (n16,), torch.bfloat16) for p in range(W)]
        peers_s = [stage if p == rank else h_s.get_buffer(p, (stage.numel(),), torch.bfloat16) for p in range(W)]
        # word-embedding rows of the window (sparse exchange): one slot per rank, filled by peer copies
        ROWS_MAX = 8192 * 1024      # elements per rank slot (8192 rows of H = 1024)
        rows = symm_mem.empty(W * ROWS_MAX, dtype=torch.bfloat16, device=dev)
        h_r = symm_mem.rendezvous(rows, group)
        peers_r = [rows if p == rank else h_r.get_buffer(p, (rows.numel(),), torch.bfloat16) for p in range(W)]
        self.g16 = g16
        self.g16_lo = g16_lo
        self.peer = dict(rank=rank, plan=plan, stage=stage, h=h_g, h_s=h_s, peers_g=peers_g, peers_s=peers_s,
                         tail=(tail_lo is not None), rows=rows, h_r=h_r, peers_r=peers_r, rows_max=ROWS_MAX,
                         ce_streams=[torch.cuda.Stream() for _ in range(3)])
        torch.cuda.synchronize()

    def _peer_reduce(self, idx):
        """Steps 1-4 of _setup_peer for bucket `idx` on the current (side) stream."""
        pr = self.peer
        W, r = self.world, pr["rank"]
        off, ln, sl, soff = pr["plan"][idx - 1] if idx >= 1 else pr["plan"][-1]   # idx 0 = the embedding tail

        def sl_range(k):
            a = min(k * sl, ln)
            return a, min(a + sl, ln)

        # 1. my copy of slice p -> peer p's staging slot r
        for d in range(1, W):
            p = (r + d) % W          # staggered targets: no two ranks hit the same peer at the same time
            a, b = sl_range(p)
            if b > a:
                pr["peers_s"][p][soff + r * sl: soff + r * sl + (b - a)].copy_(self.g16[off + a: off + b], non_blocking=True)
        pr["h"].barrier(channel=0)
        # 2. reduce my slice (own value first, then peers in rank order)
        a, b = sl_range(r)
        if b > a:
            n = b - a
            n8 = (n + 7) // 8 * 8     # slices start 16-byte aligned; a ragged tail only exists on the last slice
            srcs = [pr["stage"][soff + p * sl: soff + p * sl + n8] for p in range(W) if p != r]
            arr = (C.c_void_p * len(srcs))(*[t.data_ptr() for t in srcs])
            ops._call("b200u_slice_sum_bf16", P(self.g16[off + a: off + a + n8]), arr, len(srcs), C.c_size_t(n8))
            # 3. reduced slice -> every peer's bucket
            for d in range(1, W):
                p = (r + d) % W
                pr["peers_g"][p][off + a: off + b].copy_(self.g16[off + a: off + b], non_blocking=True)
        pr["h"].barrier(channel=0)

    def bf16_range(self):
        """(buffer, lo, hi) of the gradient range the optimizer must read as bf16, or (None, 0, 0)."""
        if self.g16 is None:
            return None, 0, 0
        return self.g16, self.g16_lo, self.g16_lo + self.g16.numel()

    def reduce_bucket(self, idx):
        lo, hi = self.segments[idx]
        self.reduced.append(idx)
        if hi <= lo or self.world <= 1:
            return
        buf = self.grad[lo:hi]
        if self.g16 is not None and idx >= 1:
            buf = self.g16[lo - self.g16_lo:hi - self.g16_lo]
            if self.sync or not self.grad.is_cuda:
                ops.cast_f32_to_bf16(self.grad[lo:hi], buf)
                if self.peer is not None:
                    self._peer_reduce(idx)
                    return
            else:
                # the fp32 -> bf16 cast of the bucket leaves the backward's critical stream: it runs on a side
                # stream forked here, and the all-reduce (NCCL stream) is ordered behind it
                cur = torch.cuda.current_stream()
                if self._cast_stream is None:
                    self._cast_stream = torch.cuda.Stream()
                self._cast_stream.wait_stream(cur)
                with torch.cuda.stream(self._cast_stream):
                    ops.cast_f32_to_bf16(self.grad[lo:hi], buf)
                    if self.peer is not None:
                        self._peer_reduce(idx)
                        self._side_work = True
                    else:
                        self.pending.append(torch.distributed.all_reduce(buf, group=self.pg, async_op=True))
                return
        if self.sync:
            torch.distributed.all_reduce(buf, group=self.pg)
        else:
            self.pending.append(torch.distributed.all_reduce(buf, group=self.pg, async_op=True))

    def reduce_tail(self):
        """Dense rest of the embedding bucket (position / type tables, LayerNorms, image embedder): cast to bf16
        into the symmetric range and exchanged like a layer bucket, on the side stream."""
        pr = self.peer
        lo = self.g16_lo
        hi = self.segments[0][1]
        cur = torch.cuda.current_stream()
        if self._cast_stream is None:
            self._cast_stream = torch.cuda.Stream()
        self._cast_stream.wait_stream(cur)
        with torch.cuda.stream(self._cast_stream):
            ops.cast_f32_to_bf16(self.grad[lo:hi], self.g16[0:hi - lo])
            self._peer_reduce(0)
        self._side_work = True

    def allgather_rows(self, rows):
        """All ranks' [n, H] bf16 row blocks -> [world * n, H] (rank-major), moved by the copy engines: every rank
        copies its block into its slot of every peer's symmetric buffer (peer copies spread over helper streams so
        several engines / NVLink paths run at once), then one barrier. None when the block does not fit."""
        pr = self.peer
        n, H = rows.shape
        cnt = n * H
        if pr is None or cnt > pr["rows_max"] or cnt % 8:
            return None
        W, r = self.world, pr["rank"]
        cur = torch.cuda.current_stream()
        src = rows.reshape(-1)
        ev = cur.record_event()
        streams = pr["ce_streams"]
        for i, st in enumerate(streams):
            st.wait_event(ev)
        for d in range(1, W):
            p = (r + d) % W
            with torch.cuda.stream(streams[(d - 1) % len(streams)]):
                pr["peers_r"][p][r * cnt:(r + 1) * cnt].copy_(src, non_blocking=True)
        pr["rows"][r * cnt:(r + 1) * cnt].copy_(src, non_blocking=True)
        for st in streams[:min(len(streams), W - 1)]:
            cur.wait_stream(st)
        rows.record_stream(cur)
        pr["h_r"].barrier(channel=0)
        return pr["rows"][:W * cnt].view(W * n, H)

    def join_side(self):
        """Order the current stream behind the side stream's bucket exchanges issued so far."""
        if self._side_work and self._cast_stream is not None:
            torch.cuda.current_stream().wait_stream(self._cast_stream)
            self._side_work = False

    def wait(self):
        self.join_side()
        for w in self.pending:
            w.wait()
        self.pending = []
        order, self.reduced = self.reduced, []
        return order


class TrainStep(object):
    def __init__(self, model, lr=3e-5, weight_decay=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 gradient_accumulation=2, max_grad_norm=5.0, pos_wt=1.8, process_group=None,
                 overlap_comm=True, comm_sm_reserve=0, fuse_window=False, comm_dtype=torch.bfloat16,
                 frozen=(), data_parallel=True, comm_impl="auto"):
        self.model = model
        self.um = model.uniter_model if hasattr(model, "uniter_model") else model.uniter
        self.accum = int(gradient_accumulation)
        self.max_grad_norm = float(max_grad_norm)
        self.pos_wt = float(pos_wt)
        self.betas, self.eps = betas, eps
        self.pg = process_group
        self.world = 1
        # data_parallel=False: a single-replica step even when a process group exists (reference runs in tests)
        if data_parallel and (process_group is not None or
                              (torch.distributed.is_available() and torch.distributed.is_initialized())):
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
        self._ids_stream = None
        self._ids_event = None
        self._sumsq_done_from = None
        self._word_sq = None
        self._ce_rows = os.environ.get("B200U_DP_ROWS", "1") != "0"
        self._sumsq_stream = None
        self._sumsq_on_side = False
        self.early_sumsq = os.environ.get("B200U_EARLY_SUMSQ", "1") != "0"

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

        # weight-decay runs over the flat layout. Parameters that get no gradient are SKIPPED, like
        # torch.optim.Adam skips tensors whose .grad is None (the reference's fine-tuning never touches
        # img_embeddings.mask_embedding.weight): `frozen` names / requires_grad=False up front, the rest is
        # detected once from the first window's gradients (see optimizer_step).
        self.weight_decay = float(weight_decay)
        self.skip = set(n for n, p, _, _ in store.entries if (not p.requires_grad) or any(f in n for f in frozen))
        self._skip_detected = False
        self._build_runs()

        # Word-embedding gradient [vocab, H] (20 % of all parameters, produced LAST by every backward, so its
        # dense all-reduce could not overlap anything, and row-sparse: <= B*T rows per micro-batch): with
        # world_size > 1 the micro-batches hand their touched rows over instead of scattering them, and after
        # the last backward the rows of the whole window are all-gathered and applied by every rank with a
        # deterministic segment add (sparse_word). No dense all-reduce of the table at all.
        self.word_slice = None
        for name, p, off, cnt in store.entries:
            if name.endswith("embeddings.word_embeddings.weight"):
                self.word_slice = (off, off + cnt, p)
        # gradient buckets: index 0 = embeddings, 1.. = encoder layers (the last also holds pooler + head)
        tail_lo = self.word_slice[1] if (self.word_slice is not None and self.word_slice[0] == 0 and overlap_comm
                                         and os.environ.get("B200U_DP_TAIL", "0") == "1") else None
        # (B200U_DP_TAIL=1 also moves the dense rest of the embedding bucket onto the copy engines as one more bf16
        #  bucket: measured -15 us on 2 GPUs but +70 us on 8, where its two extra barriers sit on the exposed tail, so
        #  the fp32 NCCL all-reduce beside the row exchange stays the default)
        self.comm = GradBuckets([(e[0], e[2]) for e in store.entries], store.grad, process_group, self.world,
                                comm_dtype=comm_dtype, comm_impl=comm_impl, tail_lo=tail_lo)
        self.buckets = self.comm.segments
        self.comm.sync = not overlap_comm
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

    def _on_layer_done_local(self, layer_idx):
        """Single replica: the layer's gradient range is final once its backward is enqueued, so its share of
        the clipping norm is taken NOW on a side stream (an HBM-bound read that co-resides with the next
        layer's tensor-core GEMMs) instead of as one 340 MB pass in front of the optimizer; optimizer_step
        only adds the embedding range. Layers finish in descending order, so the summed range stays one
        suffix [_sumsq_done_from, n) of the flat gradient."""
        lo, hi = self.buckets[layer_idx + 1]
        cur = torch.cuda.current_stream()
        if self._sumsq_stream is None:
            self._sumsq_stream = torch.cuda.Stream()
        if self._sumsq_done_from is None:
            self.sumsq.zero_()
            if hi != self.store.grad.numel():
                return      # not the last bucket first: keep the one-pass path
        elif hi != self._sumsq_done_from:
            return
        s = self._sumsq_stream
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            ops._call("b200u_grad_sumsq", P(self.store.grad[lo:hi]), C.c_size_t(hi - lo), P(self.sumsq), None,
                      C.c_size_t(0), C.c_size_t(0))
        self._sumsq_done_from = lo
        self._sumsq_on_side = True

    def _layer_cb(self, final):
        """Backward hook for the encoder layers of a pass whose gradients are final (`final`): bucket all-reduce
        in data parallel, early norm share on a single replica."""
        if not final:
            return None
        if self.world > 1:
            return self._on_layer_done if self.overlap_comm else None
        return self._on_layer_done_local if (self.early_sumsq and self.store.grad.is_cuda) else None

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
        self.um._layer_grad_ready_cb = self._layer_cb(last)
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

    def _start_word_exchange(self, batches):
        """Window start (data parallel): the token ids of the window are known BEFORE any compute, so their
        exchange and the sort that orders the word-embedding rows for the deterministic segment add run on a
        side stream beside the forward pass; after the last backward only the rows themselves travel."""
        if not (self.world > 1 and self.overlap_comm and self.sparse_word and self.word_slice is not None):
            return
        dist = torch.distributed
        cur = torch.cuda.current_stream()
        if self._ids_stream is None:
            self._ids_stream = torch.cuda.Stream()
        s = self._ids_stream
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            ids = torch.cat([b["input_ids"].reshape(-1) for b in batches]).contiguous()
            all_ids = torch.empty(self.world * ids.numel(), device=ids.device, dtype=ids.dtype)
            dist.all_gather_into_tensor(all_ids, ids, group=self.pg)
            self._ids_sorted, self._ids_perm = torch.sort(all_ids, stable=True)
            self._ids_event = s.record_event()
        for t in (self._ids_sorted, self._ids_perm):
            t.record_stream(cur)

    def _finish_sparse_word(self):
        """Embedding bucket without the word table (dense, small) + the sparse word-row exchange."""
        dist = torch.distributed
        lo, hi, wp = self.word_slice
        b_lo, b_hi = self.buckets[0]
        self.comm.reduced.append(0)
        pad = self._word_rows[0][2]
        cur = torch.cuda.current_stream()
        for r, i, _ in self._word_rows:      # earlier micro-batches produced their rows on the other stream
            r.record_stream(cur)
            i.record_stream(cur)
        rows = self._word_rows[0][0] if len(self._word_rows) == 1 else torch.cat([r for r, _, _ in self._word_rows], 0)
        self._word_rows = None
        n, H = rows.shape
        rows = rows.contiguous()
        peer_tail = self.comm.peer is not None and self.comm.peer["tail"]
        if peer_tail:
            # dense rest of the embedding bucket: bf16 two-shot exchange on the side stream, beside the row copies
            self.comm.reduce_tail()
            self.comm._tail_done = True
        all_rows = self.comm.allgather_rows(rows) if (self.comm.peer is not None and self._ce_rows) else None
        layer_handles, self.comm.pending = self.comm.pending, []
        gather = None
        if all_rows is None:
            all_rows = torch.empty(self.world * n, H, device=rows.device, dtype=rows.dtype)
            gather = dist.all_gather_into_tensor(all_rows, rows, group=self.pg, async_op=True)
        if not peer_tail:
            # the dense remainder of the embedding bucket (position / type tables, LayerNorms, image embedder)
            # is reduced behind the row exchange, beside the segment add below
            for a, b in ((b_lo, min(lo, b_hi)), (max(hi, b_lo), b_hi)):
                if b > a:
                    self.comm.pending.append(dist.all_reduce(self.store.grad[a:b], group=self.pg, async_op=True))
        # while those travel: squared norm of the (already reduced) bf16 range
        self._sumsq_layers_early(layer_handles)
        if gather is not None:
            gather.wait()
        # identical (all_rows, ids) on every rank + a deterministic, atomic-free segment add (rows sorted by
        # id by the permutation computed at the window start, each run summed in order) => bit-identical
        # word-embedding gradients on all replicas
        cur.wait_event(self._ids_event)
        tot = self.world * n
        assert self._ids_sorted.numel() == tot, "word-row exchange: ids of the window do not match its rows"
        # the table was zero before (the optimizer clears the gradients): the segment add also takes the table's
        # share of the clipping norm from the rows it writes, so nobody reads the 89 MB table for it
        fuse_norm = self._sumsq_done_from is not None and self._sumsq_done_from == hi and lo == 0
        if fuse_norm:
            if self._word_sq is None:
                self._word_sq = torch.zeros(256, device=rows.device, dtype=torch.float64)
            self._word_sq.zero_()
        ops._call("b200u_embedding_segment_add", P(all_rows), P(self._ids_sorted), P(self._ids_perm),
                  P(self.store.grad[lo:hi]), tot, H, C.c_longlong(pad), C.c_longlong((hi - lo) // H),
                  P(self._word_sq) if fuse_norm else None, 256)
        if fuse_norm:
            self.sumsq.add_(self._word_sq.sum())
            self._sumsq_done_from = 0

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
        self.um._layer_grad_ready_cb = self._layer_cb(True)
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
    def _build_runs(self):
        """Run table of b200u_adam_step over the flat layout: constant weight decay per run, -1 = skip."""
        store, dev = self.store, self.dev
        n = store.flat.numel()
        starts, wds = [], []
        for name, p, off, cnt in store.entries:
            if name in self.skip:
                wd = -1.0
            else:
                wd = 0.0 if any(nd in name for nd in NO_DECAY) else self.weight_decay
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

    def _detect_untouched(self):
        """First optimizer step: parameters no backward kernel accumulated a gradient into during the first
        window (unused branches such as img_embeddings.mask_embedding in fine-tuning) join the skip set,
        matching torch.optim.Adam's `grad is None` behaviour. Host-side bookkeeping only (the Functions
        record their gradient targets in FlatStore.touched), so it also works under CUDA-graph capture and
        takes the same decision on every data-parallel rank."""
        self._skip_detected = True
        new = set(name for name, p, _, _ in self.store.entries if id(p) not in self.store.touched)
        if not new.issubset(self.skip):
            self.skip |= new
            self._build_runs()

    def _sumsq_layers_early(self, layer_handles):
        """Data parallel, sparse word exchange: the squared norm of the encoder-layer range (75 % of the
        gradient) is taken as soon as its buckets are reduced, beside the collectives of the embedding tail,
        instead of after them; optimizer_step then only adds the embedding range."""
        for h in layer_handles:
            h.wait()
        self.comm.join_side()
        g16, lo16, hi16 = self.comm.bf16_range()
        if g16 is None:
            return
        g = self.store.grad
        self.sumsq.zero_()
        # elements [lo16, hi16): read from the bf16 buffer; the call covers exactly that range
        ops._call("b200u_grad_sumsq", P(g[lo16:hi16]), C.c_size_t(hi16 - lo16), P(self.sumsq), P(g16),
                  C.c_size_t(0), C.c_size_t(hi16 - lo16))
        self._sumsq_done_from = lo16

    def optimizer_step(self):
        self.comm.wait()
        if not self._skip_detected:
            self._detect_untouched()
        g = self.store.grad
        n = g.numel()
        g16, lo16, hi16 = self.comm.bf16_range()
        if self.comm.peer is not None and self.comm.peer["tail"]:
            if not getattr(self.comm, "_tail_done", False):
                # the embedding bucket travelled densely in fp32 this step: refresh its bf16 copy, which Adam reads
                b_hi = self.buckets[0][1]
                ops.cast_f32_to_bf16(g[lo16:b_hi], g16[0:b_hi - lo16])
            self.comm._tail_done = False
        if self._sumsq_done_from is not None:
            # the layer range was summed early: add the embedding range [0, lo16) (fp32)
            if self._sumsq_on_side:
                torch.cuda.current_stream().wait_stream(self._sumsq_stream)
                self._sumsq_on_side = False
            if self._sumsq_done_from > 0:
                ops._call("b200u_grad_sumsq", P(g), C.c_size_t(self._sumsq_done_from), P(self.sumsq), None,
                          C.c_size_t(0), C.c_size_t(0))
            self._sumsq_done_from = None
        else:
            self.sumsq.zero_()
            ops._call("b200u_grad_sumsq", P(g), C.c_size_t(n), P(self.sumsq), P(g16), C.c_size_t(lo16), C.c_size_t(hi16))
        pre = 1.0 / (self.accum * self.world)  # average_gradients + mean over ranks
        ops._call("b200u_clip_coef", P(self.sumsq), pre, self.max_grad_norm, P(self.coef), P(self.gnorm))
        ops.counter_add(self.step_t, 1)
        ops._call("b200u_adam_step", P(self.store.flat), P(g), P(self.m), P(self.v), P(self.store.shadow),
                  C.c_size_t(n), P(self.run_start), P(self.run_wd), P(self.chunk_run), self.num_runs,
                  P(self.coef), P(self.lr_t), P(self.step_t), self.betas[0], self.betas[1], self.eps, 1,
                  P(g16), C.c_size_t(lo16), C.c_size_t(hi16))

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        """Optimizer state in torch.optim.Adam's layout ('state': {index: step / exp_avg / exp_avg_sq} in
        named_parameters() order, 'param_groups'), what the reference saves as `optimizer_state_dict`
        (utils/save.py:57-64) so a run can be resumed by either implementation."""
        step = int(self.step_t.item())
        F_.check_input_errors()   # a checkpoint of a run that consumed invalid indices must not be written silently
        state = {}
        for i, (name, p) in enumerate(self.model.named_parameters()):
            off, cnt = self.store.index[id(p)]
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": self.m[off:off + cnt].view(p.shape).clone(),
                        "exp_avg_sq": self.v[off:off + cnt].view(p.shape).clone()}
        names = [n for n, _ in self.model.named_parameters()]
        return {"state": state,
                "param_groups": [{"lr": float(self.lr_t.item()), "betas": tuple(self.betas), "eps": self.eps,
                                  "weight_decay": self.weight_decay, "params": list(range(len(names)))}],
                "b200u": {"names": names, "host_step": self.host_step, "skip": sorted(self.skip)}}

    def load_state_dict(self, sd):
        params = list(self.model.named_parameters())
        for i, (name, p) in enumerate(params):
            st = sd["state"].get(i)
            if st is None:
                continue
            off, cnt = self.store.index[id(p)]
            self.m[off:off + cnt].copy_(st["exp_avg"].reshape(-1))
            self.v[off:off + cnt].copy_(st["exp_avg_sq"].reshape(-1))
        steps = [float(v["step"]) for v in sd["state"].values()]
        self.step_t.fill_(int(max(steps)) if steps else 0)
        if sd.get("param_groups"):
            self.lr_t.fill_(float(sd["param_groups"][0]["lr"]))
        extra = sd.get("b200u", {})
        self.host_step = int(extra.get("host_step", self.host_step))
        if "skip" in extra:
            self.skip = set(extra["skip"])
            self._skip_detected = True
            self._build_runs()
        self.store.refresh_shadow(force=True)

    def set_lr(self, lr):
        self.lr_t.fill_(lr)

    # ------------------------------------------------------------------ public: one optimizer step
    def step(self, batches, optimizer=True):
        """batches: list of `gradient_accumulation` device batch dicts. Returns the list of
        (loss[1], probs[B]) device tensors of the micro-batches. optimizer=False (single replica, measurement):
        forward + backward of the window only, then the gradients are cleared.

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
        self._start_word_exchange(batches)
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
        if optimizer:
            self.optimizer_step()
        else:
            assert self.world == 1, "optimizer=False is a single-replica measurement mode"
            if self._sumsq_on_side:
                torch.cuda.current_stream().wait_stream(self._sumsq_stream)
                self._sumsq_on_side = False
            self._sumsq_done_from = None
            self.store.grad.zero_()
        self.host_step += 1
        return outs

    # ------------------------------------------------------------------ CUDA-graph replay
    def capture(self, example_batches, warmup=3, optimizer=True):
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
                self.step(static, optimizer=optimizer)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        mode = "thread_local" if self.world > 1 else "global"
        if not self._skip_detected:
            # no eager step has run yet: a throw-away capture (nothing executes, no state changes) records
            # which parameters receive gradients, so the optimizer's skip runs are final before the real one
            self._skip_detected = True
            hs = self.host_step
            dry = torch.cuda.CUDAGraph()
            with torch.cuda.graph(dry, capture_error_mode=mode):
                self.step(static, optimizer=optimizer)
            del dry
            self.host_step = hs
            self._detect_untouched()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode=mode):
            outs = self.step(static, optimizer=optimizer)
        self._graph, self._static, self._static_out = graph, static, outs
        return static

    def load_static(self, batches):
        for dst, src in zip(self._static, batches):
            for k, v in dst.items():
                v.copy_(src[k], non_blocking=True)

    def replay(self):
        self.store.refresh_shadow()   # cheap when nothing changed; picks up load_state_dict / manual edits
        self._graph.replay()
        self.host_step += 1
        return self._static_out


class PretrainStep(TrainStep):
    """One optimizer step per task batch for `UniterForPretraining` (BASELINE config 5): MLM, MRFR and ITM with
    the IPOT word-region alignment, round robin — the multi-task driver the reference only sketches
    (`MetaLoader`, data/pretrain_meme_dataset.py:21-58, is never used by a trainer). The task losses are the
    per-element losses `UniterForPretraining.forward(batch, task)` returns (model/pretrain.py:65-233), averaged;
    the optimizer is the same fused Adam / clip step as fine-tuning with a window of one batch.

    The word-embedding table is also the tied MLM decoder weight (model/layer.py:204-221), whose gradient is
    dense, so the data-parallel step all-reduces the embedding bucket densely instead of exchanging rows."""

    def __init__(self, model, tasks=("mlm", "mrfr", "itm"), **kw):
        kw.setdefault("gradient_accumulation", 1)
        super().__init__(model, **kw)
        self.tasks = tuple(tasks)
        self.sparse_word = False
        self._task_i = 0

    def task_step(self, batch, task=None):
        """Forward + backward + optimizer step of one task batch. Returns (mean loss [1] device tensor, task)."""
        if task is None:
            task = self.tasks[self._task_i % len(self.tasks)]
            self._task_i += 1
        comm = self.world > 1
        self.um._layer_grad_ready_cb = self._on_layer_done if (comm and self.overlap_comm) else None
        self.um._sparse_word_cb = None
        try:
            loss_vec = self.model(batch, task, compute_loss=True)
        finally:
            self.um._layer_grad_ready_cb = None
        loss = loss_vec.float().mean()
        loss.backward()
        if comm:
            if not self.overlap_comm:
                for i in range(len(self.buckets) - 1, 0, -1):
                    self._allreduce_bucket(i)
            self._allreduce_bucket(0)
        self.optimizer_step()
        self.host_step += 1
        return loss.detach(), task
