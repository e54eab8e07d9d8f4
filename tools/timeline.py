"""Kernel timeline of one captured optimizer step (CUPTI through torch.profiler): every kernel of two graph
replays with its stream, start and duration -> gpurun_out/<tag>_timeline.csv + a per-kernel summary.

    python tools/timeline.py [tag] [--eager] [--large]
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from meme_challenge_b200.data.synthetic import synth_batch  # noqa: E402
from meme_challenge_b200.model.meme_uniter import MemeUniter  # noqa: E402
from meme_challenge_b200.model.model import UniterConfig, UniterModel  # noqa: E402
from meme_challenge_b200.train import TrainStep  # noqa: E402

tag = next((a for a in sys.argv[1:] if not a.startswith("--")), "step")
BASE = dict(vocab_size=28996, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
            intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
            attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
            initializer_range=0.02)
if "--large" in sys.argv:
    BASE.update(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)
rank, world, lrank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
if world > 1:
    os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("TL_NCCL_CTAS", "16"))
    torch.distributed.init_process_group("nccl", device_id=dev)
if "--no-pdl" in sys.argv:
    # without programmatic dependent launch a kernel's CUPTI duration is its own run time (with PDL it includes the
    # wait for its predecessor): use this mode to read per-kernel times, the default to read the real schedule
    from meme_challenge_b200 import _lib
    _lib.lib().b200u_set_pdl(0)
torch.manual_seed(0)
cfg = UniterConfig.from_dict(BASE)
model = MemeUniter(UniterModel(cfg, 2048), BASE["hidden_size"], 1).to(dev).train()
ts = TrainStep(model, gradient_accumulation=2, fuse_window="--fused" in sys.argv)
bs = []
for i in range(2):
    b = synth_batch(16, 64, 100, seed=1234 + i + 100 * rank)
    b = {k: v.to(dev) for k, v in b.items() if torch.is_tensor(v)}
    b["labels"] = b["labels"].float()
    bs.append(b)
eager = "--eager" in sys.argv
if not eager:
    ts.capture(bs, warmup=2)
    run = ts.replay
else:
    run = lambda: ts.step(bs)
for _ in range(3):
    run()
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
if rank == 0:
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            run()
        torch.cuda.synchronize()
else:
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    torch.distributed.barrier()
    os._exit(0)
os.makedirs("gpurun_out", exist_ok=True)
path = "gpurun_out/%s_trace.json" % tag
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ks.sort(key=lambda e: e["ts"])
t0 = ks[0]["ts"]
with open("gpurun_out/%s_timeline.csv" % tag, "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for e in ks:
        f.write("%.3f,%.3f,%s,%s\n" % (e["ts"] - t0, e["dur"], e.get("args", {}).get("stream", "?"),
                                      e["name"].replace(",", ";")[:120]))
os.remove(path)
print("kernels:", len(ks), "span %.1f us" % (ks[-1]["ts"] + ks[-1]["dur"] - t0))
if world > 1:
    torch.distributed.barrier()
    os._exit(0)
