"""Headline benchmark: UNITER-base fine-tune fwd+bwd memes/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one optimizer step of the reference recipe (README.md:60, train_template.py:89-109)
on BASELINE config 2: gradient_accumulation = 2 micro-batches of 16 memes (100 regions x 2048-d
+ 7-d boxes, 64 tokens, joint length 164), each forward + backward with dropout 0.1 and the
pos_wt 1.8 BCE loss, then grad averaging, clip_grad_norm_(5), Adam(L2 1e-3) and zero_grad.
Synthetic data, random-init weights. `value` times K steps with the inputs already
resident in HBM (CUDA-graph replay of the whole step); `e2e` times the same steps through the public
TrainStep API fed from pinned HOST buffers: the H2D copy of step i+1 overlaps step i on a copy stream,
every step refreshes the graph's inputs and reads its loss back (D2H). `roofline` times every GEMM
shape of the step live (CUDA events around captured back-to-back launches) against the measured
sustained bf16 peak; `cpu_baseline` times the oracle port of the reference's CPU path on the host cores.
For N > 1 run under torchrun (one rank per GPU, NCCL): each rank processes its own 16-meme
micro-batches (weak scaling); the per-layer gradient all-reduces and the sparse word-embedding row
exchange are captured in the same CUDA graph and overlap the last backward pass.
`--impl reference` times the reference's CPU implementation (oracle port) on the host cores.
Only the result line is written to stdout.
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
B, T, R, ACCUM = 16, 64, 100, 2
GFLOP_PER_MEME = 87.2  # SURVEY.md §8d / BASELINE.md §4: algorithmic fwd+bwd FLOPs, base, C2
METRIC = "UNITER-base fwd+bwd memes/s"
UNIT = "memes/s"
WORKLOAD = ("C2: UNITER-base fine-tune step = 2 micro-batches x 16 memes fwd+bwd (100 regions x 2048-d + 7-d box, "
            "64 tokens, L=164, dropout 0.1, pos_wt 1.8 BCE) + grad-average + clip 5 + Adam(L2 1e-3)")


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
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def _cpu_fwd_bwd_memes_per_s(batch_memes, iters, warmup, threads):
    """Oracle port of the reference's CPU path (fp32, training mode incl. dropout), fwd+bwd."""
    from oracle import uniter_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    from meme_challenge_b200.model.meme_uniter import MemeUniter
    from meme_challenge_b200.model.model import UniterConfig, UniterModel
    cfg = UniterConfig.from_dict(BASE)
    m = MemeUniter(UniterModel(cfg, 2048), 768, 1)   # only used as a weight container (init_weights)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    b = O.synth_batch(batch_memes, T, R, seed=1234)
    times = []
    for it in range(warmup + iters):
        t0 = time.perf_counter()
        logits = O.meme_uniter_forward(sd, BASE, input_ids=b["input_ids"], position_ids=b["position_ids"],
                                       img_feat=b["img_feat"], img_pos_feat=b["img_pos_feat"],
                                       attention_mask=b["attn_mask"], gather_index=b["gather_index"],
                                       p_hidden=0.1, p_attn=0.1, training=True)
        loss = O.bce_loss(logits, b["labels"], 1.8)
        loss.backward()
        for v in sd.values():
            v.grad = None
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return batch_memes * len(times) / sum(times), sum(times)


def _gemm_roofline(dev):
    """Per optimizer step: (sum of GEMM durations in ms, sum of algorithmic FLOPs, launches, detail)."""
    from meme_challenge_b200 import _lib, ops
    M, H, I, D, NR = B * (T + R), BASE["hidden_size"], BASE["intermediate_size"], 2048, B * R
    E = _lib
    per_layer = [("qkv_fwd", M, 3 * H, H, 0, 0, E.EPI_STORE), ("attn_out_fwd", M, H, H, 0, 0, E.EPI_BIAS_DROP_RES),
                 ("ffn1_fwd", M, I, H, 0, 0, E.EPI_BIAS_GELU), ("ffn2_fwd", M, H, I, 0, 0, E.EPI_BIAS_DROP_RES),
                 ("ffn2_wgrad", H, I, M, 1, 1, E.EPI_ATOMIC_F32), ("ffn2_dgrad", M, I, H, 0, 1, E.EPI_DGELU),
                 ("ffn1_wgrad", I, H, M, 1, 1, E.EPI_ATOMIC_F32), ("ffn1_dgrad", M, H, I, 0, 1, E.EPI_ADD),
                 ("attn_out_wgrad", H, H, M, 1, 1, E.EPI_ATOMIC_F32), ("attn_out_dgrad", M, H, H, 0, 1, E.EPI_STORE),
                 ("qkv_wgrad", 3 * H, H, M, 1, 1, E.EPI_ATOMIC_F32), ("qkv_dgrad", M, H, 3 * H, 0, 1, E.EPI_ADD)]
    shapes = [(n, 2 * BASE["num_hidden_layers"], m, nn, k, am, bm, ep) for (n, m, nn, k, am, bm, ep) in per_layer]
    shapes += [("img_linear_fwd", 2, NR, H, D, 0, 0, E.EPI_STORE_F32), ("img_linear_wgrad", 2, H, D, NR, 1, 1, E.EPI_ATOMIC_F32)]
    tot_ms = tot_fl = 0.0
    launches = 0
    detail = {}
    reps = 20
    for (name, count, m, n, k, am, bm, ep) in shapes:
        sets = []
        for _ in range(2):
            a = torch.randn((k, m) if am else (m, k), device=dev).bfloat16()
            b = torch.randn((k, n) if bm else (n, k), device=dev).bfloat16()
            f32 = ep in (E.EPI_ATOMIC_F32, E.EPI_STORE_F32)
            kw = dict(a_mn=bool(am), b_mn=bool(bm), epilogue=ep,
                      out=torch.zeros(m, n, device=dev, dtype=torch.float32 if f32 else torch.bfloat16))
            if ep in (E.EPI_STORE, E.EPI_BIAS_DROP_RES, E.EPI_BIAS_GELU, E.EPI_STORE_F32):
                kw["bias"] = torch.randn(n, device=dev)
            if ep in (E.EPI_BIAS_DROP_RES, E.EPI_ADD, E.EPI_DGELU):
                kw["res"] = torch.randn(m, n, device=dev).bfloat16()
            if ep == E.EPI_BIAS_GELU:
                kw["out2"] = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
            sets.append((a, b, kw))
        for a, b, kw in sets:
            ops.gemm(a, b, **kw)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(reps):
                a, b, kw = sets[i % 2]
                ops.gemm(a, b, **kw)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        detail[name] = round(us, 2)
        tot_ms += count * us * 1e-3
        tot_fl += count * 2.0 * m * n * k
        launches += count
    return tot_ms, tot_fl, launches, detail


def run_reference(args, rank):
    """Reference arm: the reference's CPU implementation of the path (oracle port; the reference ships no
    buildable native code) on all host cores. Each step = fwd+bwd of a bounded sample of the C2 workload,
    sized from a probe so that warmup + steps finish within ~2.5 minutes."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    probe, _ = _cpu_fwd_bwd_memes_per_s(2, 1, 1, threads)          # memes/s on a 2-meme probe
    budget_s = 150.0
    sample_memes = int(max(1, min(B, probe * budget_s / max(1, args.steps + args.warmup))))
    v, spent = _cpu_fwd_bwd_memes_per_s(sample_memes, args.steps, args.warmup, threads)
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * spent / max(1, args.steps), 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": round(v, 3), "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "oracle port of the reference CPU path (fp32, dropout on): each step = fwd+bwd of "
                                       "%d memes of the C2 shape; optimizer excluded" % sample_memes},
            "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(time.perf_counter() - t0, 1)}
    _emit(line)


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    from meme_challenge_b200 import _lib
    from meme_challenge_b200.model.meme_uniter import MemeUniter
    from meme_challenge_b200.model.model import UniterConfig, UniterModel
    from meme_challenge_b200.train import TrainStep
    from meme_challenge_b200.data.synthetic import synth_batch

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        if args.nccl_max_ctas > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_max_ctas))
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    L = _lib.lib()

    torch.manual_seed(0)
    cfg = UniterConfig.from_dict(BASE)
    model = MemeUniter(UniterModel(cfg, 2048), 768, 1).to(dev).train()
    # data-parallel modes (N > 1): "graph-overlap" captures the step INCLUDING the hook-issued NCCL bucket
    # all-reduces (forked onto the process group's stream); "graph" captures blocking collectives on the
    # compute stream (no overlap); "eager" launches everything from Python. auto = first that captures.
    dp_modes = ["graph-overlap", "graph", "eager"] if args.dp_mode == "auto" else [args.dp_mode]
    if world == 1:
        dp_modes = ["eager"] if args.no_graph else ["graph", "eager"]
    ts = TrainStep(model, lr=3e-5, weight_decay=1e-3, gradient_accumulation=ACCUM, max_grad_norm=5.0, pos_wt=1.8,
                   overlap_comm=True, comm_sm_reserve=args.comm_sm_reserve if world > 1 else 0,
                   fuse_window=(args.window == "fused"))

    # synthetic data: a ring of distinct host batches (pinned) and their device copies
    n_sets = 4
    host, devb = [], []
    for i in range(n_sets * ACCUM):
        b = synth_batch(B, T, R, seed=1234 + rank * 1000 + i)
        hb = {k: v.pin_memory() for k, v in b.items() if torch.is_tensor(v)}
        hb["labels"] = b["labels"].float().pin_memory()
        host.append(hb)
        devb.append({k: v.to(dev, non_blocking=True) for k, v in hb.items()})
    torch.cuda.synchronize()
    h2d_bytes = ACCUM * sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- capture the whole optimizer step in a CUDA graph (falls back mode by mode on failure; all
    # ranks take the same decision: a failure on any rank moves every rank to the next mode)
    use_graph = False
    launches_per_step = None
    dp_mode = "eager"
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
            before = L.b200u_launch_count()
            ts.capture(devb[:ACCUM], warmup=2)
            launches_per_step = (L.b200u_launch_count() - before) // 3  # 2 warm-ups + 1 capture
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
        before = L.b200u_launch_count()
        ts.step(devb[:ACCUM])
        launches_per_step = L.b200u_launch_count() - before

    def one_step_resident(i):
        s = (i % n_sets) * ACCUM
        if use_graph:
            ts.load_static(devb[s:s + ACCUM])   # device->device refresh of the static inputs
            return ts.replay()
        return ts.step(devb[s:s + ACCUM])

    # End-to-end pipeline: the host->device copy of step i+1 (pinned host memory, copy stream, into a
    # device staging set) overlaps the compute of step i; each step then refreshes the graph's static
    # inputs from the staging set (device->device), replays, and reads the loss back (D2H, sync).
    copy_stream = torch.cuda.Stream()
    staging = [{k: torch.empty_like(v) for k, v in b.items()} for b in devb[:ACCUM]]

    def prefetch(i):
        s = (i % n_sets) * ACCUM
        copy_stream.wait_stream(torch.cuda.current_stream())   # the previous refresh has consumed the staging set
        with torch.cuda.stream(copy_stream):
            for dst, src in zip(staging, host[s:s + ACCUM]):
                for k, v in dst.items():
                    v.copy_(src[k], non_blocking=True)

    def run_e2e(n):
        last = 0.0
        prefetch(0)
        for i in range(n):
            torch.cuda.current_stream().wait_stream(copy_stream)   # staging holds step i's inputs
            if use_graph:
                ts.load_static(staging)
                batches = None
            else:
                batches = [{k: v.clone() for k, v in b.items()} for b in staging]
            if i + 1 < n:
                prefetch(i + 1)
            outs = ts.replay() if use_graph else ts.step(batches)
            last = float(outs[-1][0].item())        # D2H read of the step's loss
        return last

    # ---- timed region 1: inputs resident in HBM
    for i in range(args.warmup):
        one_step_resident(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        one_step_resident(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()

    # ---- timed region 2: end to end from pinned host memory
    run_e2e(min(3, args.warmup))
    barrier()
    e0.record()
    last_loss = run_e2e(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()

    # ---- roofline leg (rank 0, no collectives): every GEMM shape of the step, timed live with CUDA
    # events around a captured graph of back-to-back launches (host launch gaps excluded)
    roof = None
    if rank == 0:
        sus, burst, hbm, how = _peaks()
        tms, tfl, nlaunch, detail = _gemm_roofline(dev)
        ach = tfl / (tms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": round(ach, 1), "peak": sus, "unit": "TFLOP/s",
                "frac": round(ach / sus, 4),
                # DRAM bytes per launch (dram__bytes_read + write, ncu --set full, profiles/r01_gemm_ncu_final.txt),
                # mean over the six captured shapes: reads equal the operand (+ side-input) sizes, i.e. no re-reads
                "traffic": 17.5e6,
                "kernel": "gemm_tc_kernel (tcgen05+TMA bf16 GEMM family): %d launches per optimizer step, each shape "
                          "timed as 20 back-to-back launches in a CUDA graph; achieved = sum(2MNK) / sum(duration); "
                          "peak = sustained cuBLAS bf16 (%s)" % (nlaunch, how),
                "gemm_ms_per_step": round(tms, 3), "per_shape_us": detail}

    if rank == 0:
        memes = args.steps * ACCUM * B * world
        value = memes / (ms * 1e-3)
        e2e = memes / (ms_e2e * 1e-3)
        sus, burst, hbm, how = _peaks()
        cpu = None
        if world == 1 and not args.skip_cpu:
            threads = os.cpu_count() or 1
            v, spent = _cpu_fwd_bwd_memes_per_s(B, 6, 1, threads)
            cpu = {"value": round(v, 3), "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "oracle port of the reference CPU path (fp32, dropout on): 6 timed fwd+bwd passes over "
                             "one %d-meme micro-batch of the C2 shape (%.1f s); optimizer excluded" % (B, spent)}
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": B * world, "grad_accum": ACCUM,
                           "memes_per_step": ACCUM * B * world, "parallelism": "dp%d" % world,
                           "window": args.window,
                           "cuda_graph": use_graph, "dp_mode": dp_mode if world > 1 else None,
                           "nccl_max_ctas": args.nccl_max_ctas if world > 1 else None,
                           "comm_sm_reserve": args.comm_sm_reserve if world > 1 else None,
                           "l2": "per-step working set (~0.75 GB saved activations + 1.5 GB fp32 params/grads/Adam "
                                 "state + 0.2 GB bf16 weights) exceeds the 126 MB L2; inputs rotate over %d batch sets" % n_sets,
                           "model_tflop_per_s": round(value * GFLOP_PER_MEME / 1e3, 1),
                           "mfu_vs_sustained_bf16": round(value * GFLOP_PER_MEME / 1e3 / world / sus, 4)},
                "e2e": {"value": round(e2e, 1), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": 4, "last_loss": round(last_loss, 5)},
                "gpu_launches": int(launches_per_step) * args.steps,
                "clocks": clocks, "roofline": roof}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        _emit(line)
    if world > 1:
        # A captured CUDA graph holds NCCL kernels of this communicator: drop it before the process group,
        # and leave through os._exit so a communicator teardown that blocks cannot hang the launcher.
        torch.cuda.synchronize()
        ts._graph = None
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
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--window", default="pipelined", choices=["pipelined", "fused"],
                    help="how the micro-batches of one accumulation window are executed: pipelined = one "
                         "forward/backward per micro-batch (forward i+1 beside backward i); fused = one pass over "
                         "all accum x 16 memes (same gradients and per-micro-batch losses, see TrainStep.fuse_batches)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--dp-mode", default="auto", choices=["auto", "graph-overlap", "graph", "eager"],
                    help="N > 1: how the data-parallel step is launched (auto = first mode that captures)")
    ap.add_argument("--nccl-max-ctas", type=int, default=16,
                    help="N > 1: cap NCCL's CTAs per collective (0 = NCCL default). 16 measured best on 2 and 8 B200s: "
                         "the bucket all-reduces overlap the backward pass and every NCCL CTA owns an SM while it runs")
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
