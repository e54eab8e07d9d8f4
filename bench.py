"""Headline benchmark: UNITER fine-tune fwd+bwd memes/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config base|large|pretrain]

A "step" is one optimizer step of the reference recipe (README.md:60, train_template.py:89-109) on BASELINE
config 2: gradient_accumulation = 2 micro-batches of 16 memes (100 regions x 2048-d + 7-d boxes, 64 tokens,
joint length 164), each forward + backward with dropout 0.1 and the pos_wt 1.8 BCE loss, then grad averaging,
clip_grad_norm_(5), Adam(L2 1e-3) and zero_grad. Synthetic data, random-init weights.

  value     K steps with the inputs already resident in HBM (CUDA-graph replay of the whole step)
  e2e       the same steps through the public API fed from pinned HOST memory by data.pipeline.PinnedPrefetcher:
            the H2D copies of step i+1 run on a copy stream beside step i, every step refreshes the graph's
            inputs and reads its loss back (D2H)
  roofline  every GEMM shape of the step timed live (CUDA events around captured back-to-back launches) against
            the measured BURST bf16 peak (the kernels are timed in isolation), sustained fraction beside it;
            `hbm`: the HBM-bound kernels against the measured copy bandwidth; `ot`: IPOT launch time vs the floor
  cpu_baseline / --impl reference
            the UNMODIFIED reference modules (oracle/_ref, see oracle/build_ref.py; oracle port if absent) on the
            host cores with the SAME step definition: 2 micro-batches fwd+bwd, grad/2, clip, Adam, zero_grad

`--window fused` (default) runs the two micro-batches of a window as ONE pass over 32 memes (same per-micro-batch
losses and accumulated gradient, tests/test_gpu_parity.py::test_fused_window_equals_sequential_window);
`--window pipelined` keeps one forward/backward per micro-batch (forward i+1 beside backward i). At N = 1 the
other mode is measured too and reported under config.window_other. For N > 1 run under torchrun (one rank per
GPU, NCCL): weak scaling, bf16 gradient buckets all-reduced from the backward hooks, sparse word-row exchange,
all captured in the step's CUDA graph. Only the result line is written to stdout.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

BASE = dict(vocab_size=28996, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
            intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
            attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
            initializer_range=0.02)
LARGE = dict(BASE, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)
B, T, R, ACCUM = 16, 64, 100, 2
UNIT = "memes/s"
CONFIGS = {
    # name: (model config, GFLOP per meme fwd+bwd, metric, workload)
    "base": (BASE, 87.2, "UNITER-base fwd+bwd memes/s",
             "C2: UNITER-base fine-tune step = 2 micro-batches x 16 memes fwd+bwd (100 regions x 2048-d + 7-d box, "
             "64 tokens, L=164, dropout 0.1, pos_wt 1.8 BCE) + grad-average + clip 5 + Adam(L2 1e-3)"),
    "large": (LARGE, 305.9, "UNITER-large fwd+bwd memes/s",
              "C4: UNITER-large (24 layers, H=1024) fine-tune step = 2 micro-batches x 16 memes fwd+bwd (100 regions, "
              "64 tokens, L=164, dropout 0.1, pos_wt 1.8 BCE) + grad-average + clip 5 + Adam(L2 1e-3)"),
    "pretrain": (BASE, 87.2, "UNITER-base pretraining memes/s",
                 "C5: UNITER-base pretraining, round robin MLM / MRFR / ITM(+IPOT word-region alignment): one optimizer "
                 "step per 16-meme task batch (100 regions, 64 tokens, mask prob 0.15), Adam(L2 1e-3), clip 5"),
}

_REAL_STDOUT = None


def _guard_stdout():
    """The contract is ONE JSON line on stdout: route everything else that writes to fd 1 (NCCL's version
    banner, library chatter) to stderr and keep the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1374.8), d.get("bf16_tflops", 1621.6), d.get("hbm_gbs", 6544.3), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


def _gemm_traffic():
    """DRAM bytes per launch of the dominant GEMM shapes (dram__bytes_read + write from the committed
    `ncu --set full` export of this round's build), or None when the export is not there."""
    p = os.path.join(ROOT, "profiles", "r02_gemm_dram_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("mean_bytes_per_launch")
        except Exception:  # noqa: BLE001
            return None
    return None


class ClockSampler(threading.Thread):
    """nvidia-smi SM clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules on the host cores (oracle/_ref), same step definition
# ------------------------------------------------------------------------------------------------
class CpuReferenceStep(object):
    """One optimizer step of the reference recipe on the CPU: ACCUM micro-batches of `mb` memes through
    MemeUniter.forward + BCEWithLogitsLoss(pos_weight) + backward, then grad /= ACCUM, clip_grad_norm_(5),
    Adam (utils/optim_utils.get_optimizer: L2 decay groups), zero_grad (train_template.py:89-109)."""

    def __init__(self, cfg_dict, mb, threads):
        from meme_challenge_b200.data.synthetic import synth_batch
        torch.set_num_threads(threads)
        torch.manual_seed(0)
        self.mb = mb
        H = cfg_dict["hidden_size"]
        try:
            from oracle import ref_loader
            self.kind = "reference" if ref_loader.available() else "port"
        except Exception:  # noqa: BLE001
            self.kind = "port"
        self.crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([1.8]))
        if self.kind == "reference":
            ns = ref_loader.load()
            cfg = ns.model.UniterConfig.from_dict(cfg_dict)
            self.model = ns.meme_uniter.MemeUniter(ns.model.UniterModel(cfg, 2048), H, 1).train()
            self.opt = ns.optim_utils.get_optimizer(self.model, dict(weight_decay=1e-3, optimizer="adam", beta1=0.9,
                                                                     beta2=0.999, lr=3e-5))
            self.what = "unmodified reference modules (oracle/_ref: model/{model,layer,meme_uniter}.py, utils/optim_utils.py)"
        else:
            from oracle import uniter_oracle as O
            from meme_challenge_b200.model.meme_uniter import MemeUniter
            from meme_challenge_b200.model.model import UniterConfig, UniterModel
            self.O, self.cfg_dict = O, cfg_dict
            m = MemeUniter(UniterModel(UniterConfig.from_dict(cfg_dict), 2048), H, 1)   # weight container only
            self.sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
            self.opt = torch.optim.Adam([{"params": [v for k, v in self.sd.items() if not O.is_no_decay(k)], "weight_decay": 1e-3},
                                         {"params": [v for k, v in self.sd.items() if O.is_no_decay(k)], "weight_decay": 0.0}], lr=3e-5)
            self.what = "oracle port of the reference CPU path (oracle/_ref not built)"
        self.batches = [synth_batch(mb, T, R, seed=1234 + i) for i in range(ACCUM)]

    def step(self):
        params = list(self.model.parameters()) if self.kind == "reference" else list(self.sd.values())
        for b in self.batches:
            kw = dict(input_ids=b["input_ids"], position_ids=b["position_ids"], img_feat=b["img_feat"],
                      img_pos_feat=b["img_pos_feat"], attention_mask=b["attn_mask"], gather_index=b["gather_index"])
            if self.kind == "reference":
                logits = self.model(output_all_encoded_layers=False, **kw)
            else:
                logits = self.O.meme_uniter_forward(self.sd, self.cfg_dict, p_hidden=0.1, p_attn=0.1, training=True, **kw)
            self.crit(logits.squeeze(1), b["labels"].float()).backward()
        for p in params:
            if p.grad is not None:
                p.grad /= ACCUM
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 5)
        self.opt.step()
        self.opt.zero_grad()

    def run(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        dt = time.perf_counter() - t0
        return ACCUM * self.mb * steps / dt, dt


def run_reference(args, rank):
    """Reference arm: the reference's own CPU implementation of the path on all host cores. Each step is one
    optimizer step over a bounded sample of the workload (2 micro-batches of `mb` memes), `mb` sized from a probe
    so that warmup + steps finish within ~2.5 minutes."""
    if rank != 0:
        return
    cfg_dict, _, metric, workload = CONFIGS[args.config if args.config != "pretrain" else "base"]
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    probe = CpuReferenceStep(cfg_dict, 1, threads)
    rate, _ = probe.run(1, 1)                       # memes/s of a 2 x 1-meme step (optimizer included)
    budget_s = 150.0
    mb = int(max(1, min(B, rate * budget_s / max(1, args.steps + args.warmup) / ACCUM)))
    ref = CpuReferenceStep(cfg_dict, mb, threads) if mb != 1 else probe
    v, spent = ref.run(args.steps, args.warmup)
    line = {"impl": "reference", "metric": metric, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * spent / max(1, args.steps), 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": workload},
            "cpu_baseline": {"value": round(v, 3), "unit": UNIT, "cores": threads, "kind": ref.kind,
                             "sample": "%s, fp32, dropout on: each step = %d micro-batches x %d memes of the C2 shape "
                                       "fwd+bwd + grad-average + clip 5 + Adam(L2) + zero_grad" % (ref.what, ACCUM, mb)},
            "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(time.perf_counter() - t0, 1)}
    _emit(line)


# ------------------------------------------------------------------------------------------------
def _host_ring(rank, n_sets, pretrain=False):
    """Ring of distinct pinned host batches (n_sets windows of ACCUM micro-batches)."""
    from meme_challenge_b200.data.synthetic import synth_batch, synth_pretrain_batch
    host = []
    for i in range(n_sets * ACCUM):
        if pretrain:
            b = synth_pretrain_batch(B, T, R, seed=1234 + rank * 1000 + i, variable=False)
        else:
            b = synth_batch(B, T, R, seed=1234 + rank * 1000 + i)
        hb = {k: v.pin_memory() for k, v in b.items() if torch.is_tensor(v)}
        if "labels" in hb:
            hb["labels"] = b["labels"].float().pin_memory()
        if pretrain:
            hb["ot_inputs"] = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in b["ot_inputs"].items()}
        host.append(hb)
    return host


def _to_dev(hb, dev):
    out = {}
    for k, v in hb.items():
        if torch.is_tensor(v):
            out[k] = v.to(dev, non_blocking=True)
        elif isinstance(v, dict):
            out[k] = {kk: (vv.to(dev, non_blocking=True) if torch.is_tensor(vv) else vv) for kk, vv in v.items()}
        else:
            out[k] = v
    return out


def _measure_finetune(args, model_cfg, window, rank, world, dev, host, devb, L_):
    """Build the model + TrainStep for `window`, capture, time resident + e2e. Returns a dict of measurements."""
    import torch.distributed as dist
    from meme_challenge_b200.data.pipeline import PinnedPrefetcher
    from meme_challenge_b200.model.meme_uniter import MemeUniter
    from meme_challenge_b200.model.model import UniterConfig, UniterModel
    from meme_challenge_b200.train import TrainStep

    torch.manual_seed(0)
    cfg = UniterConfig.from_dict(model_cfg)
    model = MemeUniter(UniterModel(cfg, 2048), model_cfg["hidden_size"], 1).to(dev).train()
    n_params = sum(p.numel() for p in model.parameters())
    dp_modes = ["graph-overlap", "graph", "eager"] if args.dp_mode == "auto" else [args.dp_mode]
    if world == 1:
        dp_modes = ["eager"] if args.no_graph else ["graph", "eager"]
    ts = TrainStep(model, lr=3e-5, weight_decay=1e-3, gradient_accumulation=ACCUM, max_grad_norm=5.0, pos_wt=1.8,
                   overlap_comm=True, comm_sm_reserve=args.comm_sm_reserve if world > 1 else 0,
                   fuse_window=(window == "fused"),
                   comm_dtype=(torch.bfloat16 if args.comm_dtype == "bf16" else None), comm_impl=args.dp_impl)
    n_sets = len(devb) // ACCUM

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph, launches_per_step, dp_mode = False, None, "eager"
    for mode in dp_modes:
        dp_mode = mode
        if mode == "eager":
            ts.overlap_comm = True
            ts.comm.sync = False
            break
        ts.overlap_comm = (mode == "graph-overlap")
        ts.comm.sync = not ts.overlap_comm
        ok = 1
        try:
            before = L_.b200u_launch_count()
            ts.capture(devb[:ACCUM], warmup=2)
            launches_per_step = (L_.b200u_launch_count() - before) // 3  # 2 warm-ups + 1 capture
        except Exception as e:  # noqa: BLE001
            ok = 0
            sys.stderr.write("[rank %d] CUDA graph capture in mode %s failed: %s\n" % (rank, mode, str(e)[:300]))
            ts._graph = None
            ts.comm.pending = []
            ts.comm.reduced = []
            torch.cuda.synchronize()
        if world > 1:
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if ok:
            use_graph = True
            break
    if launches_per_step is None:
        before = L_.b200u_launch_count()
        ts.step(devb[:ACCUM])
        launches_per_step = L_.b200u_launch_count() - before

    def one_step_resident(i):
        s = (i % n_sets) * ACCUM
        if use_graph:
            ts.load_static(devb[s:s + ACCUM])   # device->device refresh of the static inputs
            return ts.replay()
        return ts.step(devb[s:s + ACCUM])

    # End-to-end: the package's prefetcher copies window i+1 from pinned host memory on its copy stream while
    # window i computes; each step refreshes the graph's static inputs from the prefetched device buffers
    # (device->device), replays, and reads the loss back (D2H, sync).
    def run_e2e(n):
        windows = (host[(i % n_sets) * ACCUM:(i % n_sets) * ACCUM + ACCUM] for i in range(n))
        feeder = PinnedPrefetcher(windows, dev, depth=2, auto_prefetch=False)
        last, nbytes = 0.0, 0
        for batches in feeder:
            nbytes = feeder.h2d_bytes
            if use_graph:
                ts.load_static(batches)
                outs = ts.replay()
            else:
                outs = ts.step([dict(b) for b in batches])
            feeder.prefetch_next()                  # next window's H2D copies, issued while this step runs
            last = float(outs[-1][0].item())        # D2H read of the step's loss
        return last, nbytes

    for i in range(args.warmup):
        one_step_resident(i)
    barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        one_step_resident(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()

    run_e2e(min(3, args.warmup))
    barrier()
    e0.record()
    last_loss, h2d_bytes = run_e2e(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    from meme_challenge_b200.functional import check_input_errors
    check_input_errors()   # no kernel of the timed steps was handed an out-of-range index
    res = dict(ms=ms, ms_e2e=ms_e2e, clocks=clocks, last_loss=last_loss, h2d_bytes=h2d_bytes, use_graph=use_graph,
               dp_mode=dp_mode, launches_per_step=int(launches_per_step), n_params=n_params,
               dp_impl=("ce" if ts.comm.peer is not None else "nccl"))
    if world > 1:
        # A captured CUDA graph holds NCCL kernels of this communicator: drop it before the process group.
        torch.cuda.synchronize()
        ts._graph = None
    del ts, model
    torch.cuda.empty_cache()
    return res


def _measure_pretrain(args, model_cfg, rank, world, dev, host, devb, L_):
    """C5: UniterForPretraining, round-robin MLM / MRFR / ITM(+OT), one optimizer step per 16-meme task batch.
    Eager launches: the MLM / MRFR heads select a data-dependent number of masked rows (model/pretrain.py:129-133)."""
    import torch.distributed as dist
    from meme_challenge_b200.model.model import UniterConfig
    from meme_challenge_b200.model.pretrain import UniterForPretraining
    from meme_challenge_b200.train import PretrainStep
    torch.manual_seed(0)
    cfg = UniterConfig.from_dict(model_cfg)
    model = UniterForPretraining(cfg, 2048, 1601).to(dev).train()
    n_params = sum(p.numel() for p in model.parameters())
    ts = PretrainStep(model, lr=3e-5, weight_decay=1e-3, max_grad_norm=5.0,
                      comm_dtype=(torch.bfloat16 if args.comm_dtype == "bf16" else None))
    nb = len(devb)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    before = L_.b200u_launch_count()
    for i in range(3):
        ts.task_step(devb[i % nb])
    launches_per_step = (L_.b200u_launch_count() - before) // 3
    for i in range(max(0, args.warmup - 3)):
        ts.task_step(devb[i % nb])
    barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        ts.task_step(devb[i % nb])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    # end to end: every step copies its task batch from pinned host memory and reads the loss back
    copy_stream = torch.cuda.Stream()
    e0.record()
    last = 0.0
    nxt = None
    for i in range(args.steps):
        if nxt is None:
            with torch.cuda.stream(copy_stream):
                nxt = _to_dev(host[i % nb], dev)
        torch.cuda.current_stream().wait_stream(copy_stream)
        cur = nxt
        if i + 1 < args.steps:
            with torch.cuda.stream(copy_stream):
                nxt = _to_dev(host[(i + 1) % nb], dev)
        loss, _ = ts.task_step(cur)
        last = float(loss.item())
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    nbytes = sum(v.numel() * v.element_size() for v in host[0].values() if torch.is_tensor(v))
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    return dict(ms=ms, ms_e2e=ms_e2e, clocks=clocks, last_loss=last, h2d_bytes=nbytes, use_graph=False, dp_mode="eager",
                launches_per_step=int(launches_per_step), n_params=n_params)


def _measure_variants(args, model_cfg, dev):
    """SURVEY.md 8(d), C2 side measurements on one GPU (fused window, CUDA graphs, inputs resident):
    fwd+bwd only (no optimizer; the gradients are cleared instead) and the variable-length variant
    (txt_len ~ U{8..64}, 36-100 regions per meme: every batch set has its own padded widths and its own graph)."""
    from meme_challenge_b200.data.synthetic import synth_batch
    from meme_challenge_b200.model.meme_uniter import MemeUniter
    from meme_challenge_b200.model.model import UniterConfig, UniterModel
    from meme_challenge_b200.train import TrainStep

    def time_steps(run, n):
        for i in range(args.warmup):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def dev_batch(seed, variable):
        b = synth_batch(B, T, R, seed=seed, variable=variable)
        d = {k: v.to(dev) for k, v in b.items() if torch.is_tensor(v)}
        d["labels"] = d["labels"].float()
        return d

    out = {}
    torch.manual_seed(0)
    model = MemeUniter(UniterModel(UniterConfig.from_dict(model_cfg), 2048), model_cfg["hidden_size"], 1).to(dev).train()
    ts = TrainStep(model, lr=3e-5, weight_decay=1e-3, gradient_accumulation=ACCUM, max_grad_norm=5.0, pos_wt=1.8,
                   fuse_window=True)
    steps = max(10, args.steps // 2)
    # (a) forward + backward only
    ts.capture([dev_batch(1234 + i, False) for i in range(ACCUM)], warmup=2, optimizer=False)
    ms = time_steps(lambda i: ts.replay(), steps)
    out["fwd_bwd_only"] = {"value": round(ACCUM * B / (ms * 1e-3), 1), "ms_per_step": round(ms, 3),
                           "what": "2 x 16 memes forward + backward, gradients cleared, no clip / Adam"}
    # (b) variable-length batches, full optimizer step, one graph per batch set
    graphs, widths = [], []
    for k in range(4):
        bs = [dev_batch(5000 + ACCUM * k + i, True) for i in range(ACCUM)]
        ts.capture(bs, warmup=1)
        graphs.append((ts._graph, ts._static, ts._static_cat, ts._static_out))
        widths.append(max(int(b["attn_mask"].shape[1]) for b in bs))

    def run_var(i):
        ts._graph, ts._static, ts._static_cat, ts._static_out = graphs[i % len(graphs)]
        ts.replay()
    ms = time_steps(run_var, steps)
    out["variable_length"] = {"value": round(ACCUM * B / (ms * 1e-3), 1), "ms_per_step": round(ms, 3),
                              "what": "full optimizer step on ragged batches (txt_len ~ U{8..64}, 36-100 regions per meme; "
                                      "padded joint widths of the 4 batch sets: %s)" % widths}
    ts._graph = None
    del graphs, ts, model
    torch.cuda.empty_cache()
    return out


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    from meme_challenge_b200 import _lib, roofline

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        if args.nccl_max_ctas > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_max_ctas))
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    L_ = _lib.lib()
    model_cfg, gflop_per_meme, metric, workload = CONFIGS[args.config]
    pretrain = args.config == "pretrain"

    n_sets = 4
    host = _host_ring(rank, n_sets, pretrain)
    devb = [_to_dev(hb, dev) for hb in host]
    torch.cuda.synchronize()

    other = None
    variants = None
    if pretrain:
        main = _measure_pretrain(args, model_cfg, rank, world, dev, host, devb, L_)
        memes_per_step = B * world
    else:
        main = _measure_finetune(args, model_cfg, args.window, rank, world, dev, host, devb, L_)
        memes_per_step = ACCUM * B * world
        if world == 1 and not args.no_secondary:
            ow = "pipelined" if args.window == "fused" else "fused"
            try:
                o = _measure_finetune(args, model_cfg, ow, rank, world, dev, host, devb, L_)
                other = {"window": ow, "value": round(args.steps * memes_per_step / (o["ms"] * 1e-3), 1),
                         "ms_per_step": round(o["ms"] / args.steps, 3),
                         "e2e": round(args.steps * memes_per_step / (o["ms_e2e"] * 1e-3), 1)}
            except Exception as e:  # noqa: BLE001
                sys.stderr.write("secondary window measurement failed: %s\n" % str(e)[:300])
            if args.config == "base":
                try:
                    variants = _measure_variants(args, model_cfg, dev)
                except Exception as e:  # noqa: BLE001
                    sys.stderr.write("variant measurements failed: %s\n" % str(e)[:300])

    # ---- roofline legs (rank 0, no collectives)
    roof = None
    if rank == 0:
        sus, burst, hbm, how = _peaks()
        H, I, layers = model_cfg["hidden_size"], model_cfg["intermediate_size"], model_cfg["num_hidden_layers"]
        fused = (not pretrain) and args.window == "fused"
        passes = 1 if (fused or pretrain) else ACCUM
        Mrows = B * (T + R) * (ACCUM if fused else 1)
        tms, tfl, nlaunch, detail = roofline.gemm_family(dev, Mrows, H, I, layers, passes, B * R * (ACCUM if fused else 1))
        ach = tfl / (tms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": round(ach, 1), "peak": burst, "unit": "TFLOP/s",
                "frac": round(ach / burst, 4), "frac_of_sustained_peak": round(ach / sus, 4),
                "traffic": _gemm_traffic(),
                "kernel": "gemm_tc_kernel (tcgen05+TMA bf16 GEMM family, fused epilogues incl. LayerNorm / GELU' / bias-grad): "
                          "%d launches per optimizer step at M = %d rows, each shape timed as %d back-to-back launches in a "
                          "CUDA graph; achieved = sum(2MNK) / sum(duration); peak = burst cuBLAS bf16 (%s): kernels timed in "
                          "isolation" % (nlaunch, Mrows, roofline.REPS, how),
                "gemm_ms_per_step": round(tms, 3), "per_shape": detail}
        try:
            roof["hbm"] = roofline.hbm_kernels(dev, B * (ACCUM if fused else 1), T, R, T + R, H,
                                               model_cfg["num_attention_heads"], main["n_params"], hbm)
            roof["hbm_peak_gbs"] = hbm
            roof["ot"] = roofline.ot_kernels(dev)
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("hbm / ot roofline leg failed: %s\n" % str(e)[:300])

    if rank == 0:
        ms, ms_e2e = main["ms"], main["ms_e2e"]
        memes = args.steps * memes_per_step
        value = memes / (ms * 1e-3)
        e2e = memes / (ms_e2e * 1e-3)
        sus, burst, hbm, how = _peaks()
        cpu = None
        if world == 1 and not args.skip_cpu:
            threads = os.cpu_count() or 1
            small = args.config != "base"
            ref = CpuReferenceStep(BASE if pretrain else model_cfg, 4 if small else B, threads)
            v, spent = ref.run(1 if small else 2, 1)
            cpu = {"value": round(v, 3), "unit": UNIT, "cores": threads, "kind": ref.kind,
                   "sample": "%s, fp32, dropout on: %d timed optimizer step(s) of %d micro-batches x %d memes of the C2 shape "
                             "(fwd+bwd + grad-average + clip + Adam), %.1f s" % (ref.what, 1 if small else 2, ACCUM, ref.mb,
                                                                              spent)}
        line = {"metric": metric, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload, "global_batch": B * world, "grad_accum": 1 if pretrain else ACCUM,
                           "memes_per_step": memes_per_step, "parallelism": "dp%d" % world,
                           "window": None if pretrain else args.window,
                           "window_note": None if pretrain else (
                               "fused: the 2 micro-batches of an accumulation window run as ONE pass over 32 memes; the "
                               "per-micro-batch mean losses, their gradients and the /2 averaging are those of the "
                               "sequential window" if args.window == "fused" else
                               "pipelined: one forward/backward per micro-batch, forward i+1 beside backward i"),
                           "window_other": other,
                           "cuda_graph": main["use_graph"], "dp_mode": main["dp_mode"] if world > 1 else None,
                           "grad_comm_dtype": args.comm_dtype if world > 1 else None,
                           "grad_comm_impl": main.get("dp_impl") if world > 1 else None,
                           "nccl_max_ctas": args.nccl_max_ctas if world > 1 else None,
                           "comm_sm_reserve": args.comm_sm_reserve if world > 1 else None,
                           "l2": "per-step working set (saved activations + fp32 params/grads/Adam state + bf16 weights, "
                                 "> 2 GB) exceeds the 126 MB L2; inputs rotate over %d batch sets" % n_sets,
                           "model_tflop_per_s": round(value * gflop_per_meme / 1e3, 1),
                           "mfu_vs_sustained_bf16": round(value * gflop_per_meme / 1e3 / world / sus, 4)},
                "e2e": {"value": round(e2e, 1), "unit": UNIT, "h2d_bytes_per_step": int(main["h2d_bytes"]),
                        "d2h_bytes_per_step": 4, "last_loss": round(main["last_loss"], 5)},
                "gpu_launches": int(main["launches_per_step"]) * args.steps,
                "clocks": main["clocks"], "roofline": roof}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if variants is not None:
            line["variants"] = variants
        _emit(line)
    if world > 1:
        # leave through os._exit so a communicator teardown that blocks cannot hang the launcher
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="base", choices=sorted(CONFIGS),
                    help="base = BASELINE config 2/3 (headline), large = config 4 (24 x H=1024), pretrain = config 5")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--window", default="fused", choices=["pipelined", "fused"],
                    help="how the micro-batches of one accumulation window are executed: fused = one pass over all "
                         "accum x 16 memes (same gradients and per-micro-batch losses, see TrainStep.fuse_batches); "
                         "pipelined = one forward/backward per micro-batch (forward i+1 beside backward i)")
    ap.add_argument("--no-secondary", action="store_true", help="N = 1: do not also measure the other window mode")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--dp-mode", default="auto", choices=["auto", "graph-overlap", "graph", "eager"],
                    help="N > 1: how the data-parallel step is launched (auto = first mode that captures)")
    ap.add_argument("--dp-impl", default="auto", choices=["auto", "ce", "nccl"],
                    help="N > 1, bf16 buckets: 'ce' = two-shot all-reduce with the copy engines over symmetric memory "
                         "(no collective kernels on the SMs), 'nccl' = NCCL ring all-reduce, 'auto' = ce where available")
    ap.add_argument("--comm-dtype", default="bf16", choices=["bf16", "fp32"],
                    help="N > 1: dtype of the encoder-layer gradient buckets on the wire (fp32 accumulation stays local)")
    ap.add_argument("--nccl-max-ctas", type=int, default=16,
                    help="N > 1: cap NCCL's CTAs per collective (0 = NCCL default): the bucket all-reduces overlap the "
                         "backward pass and every NCCL CTA owns an SM while it runs")
    ap.add_argument("--comm-sm-reserve", type=int, default=0,
                    help="N > 1: SMs left to NCCL while all-reduces overlap the backward pass")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        _guard_stdout()
        run_reference(args, rank)
        return
    if args.gpus > 1 and world == 1:
        # convenience: re-launch ourselves under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    _guard_stdout()
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
