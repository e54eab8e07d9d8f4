"""torchrun target: data-parallel TrainStep correctness on N GPUs (tiny model, dropout off).
Checks (1) every rank ends with identical parameters, (2) the sparse word-embedding row exchange gives the
same parameters as the dense all-reduce path, (3) both equal a single-process run over the union of the
ranks' micro-batches with gradients averaged the same way (world * accum micro-batches per step)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
from meme_challenge_b200.model.meme_uniter import MemeUniter
from meme_challenge_b200.model.model import UniterConfig, UniterModel
from meme_challenge_b200.train import TrainStep
from meme_challenge_b200.data.synthetic import synth_batch

TINY = dict(vocab_size=120, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
            intermediate_size=256, hidden_act="gelu", hidden_dropout_prob=0.1,
            attention_probs_dropout_prob=0.1, max_position_embeddings=64, type_vocab_size=2,
            initializer_range=0.02)
IMG_DIM = 64

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)
cfg = dict(TINY); cfg["hidden_dropout_prob"] = 0.0; cfg["attention_probs_dropout_prob"] = 0.0


def batch(seed):
    b = synth_batch(4, 12, 10, seed=seed, img_dim=IMG_DIM, vocab=TINY["vocab_size"], min_txt=2, min_bb=2)
    d = {k: v.to(dev) for k, v in b.items() if torch.is_tensor(v)}
    d["labels"] = b["labels"].float().to(dev)
    return d


def run(sparse, graph, steps=2, comm_dtype=None, fused=False, impl="auto"):
    torch.manual_seed(0)
    m = MemeUniter(UniterModel(UniterConfig.from_dict(cfg), IMG_DIM), cfg["hidden_size"], 1).to(dev).train()
    ts = TrainStep(m, lr=1e-3, weight_decay=1e-3, gradient_accumulation=2, max_grad_norm=5.0, pos_wt=1.8,
                   comm_dtype=comm_dtype, fuse_window=fused, comm_impl=impl)
    ts.sparse_word = sparse
    used.append("ce" if ts.comm.peer is not None else "nccl")
    if graph:
        ts.capture([batch(100 + rank * 10), batch(101 + rank * 10)], warmup=0)
    for s in range(steps):
        bs = [batch(100 + rank * 10 + 2 * s), batch(101 + rank * 10 + 2 * s)]
        if graph:
            ts.load_static(bs); ts.replay()
        else:
            ts.step(bs)
    torch.cuda.synchronize()
    return torch.cat([p.detach().flatten() for p in m.parameters()])


def single(steps=2):
    """one process, the union of all ranks' micro-batches: grads summed, / (accum * world), clip, Adam"""
    torch.manual_seed(0)
    m = MemeUniter(UniterModel(UniterConfig.from_dict(cfg), IMG_DIM), cfg["hidden_size"], 1).to(dev).train()
    ts = TrainStep(m, lr=1e-3, weight_decay=1e-3, gradient_accumulation=2 * world, max_grad_norm=5.0, pos_wt=1.8,
                   data_parallel=False)
    for s in range(steps):
        bs = []
        for r in range(world):
            bs += [batch(100 + r * 10 + 2 * s), batch(101 + r * 10 + 2 * s)]
        ts.pipeline = False
        ts.step(bs)
    torch.cuda.synchronize()
    return torch.cat([p.detach().flatten() for p in m.parameters()])


ok = True
ref = single() if rank == 0 else None
# fp32 gradient all-reduce: every variant must land on the single-process parameters (fp32 summation order
# is the only difference); bf16 buckets: replicas still bit-identical, parameters within Adam's sensitivity
# to a 2^-9 relative perturbation of the gradients (a near-zero gradient can flip the sign of a ~lr step)
# bf16 buckets travel either with the copy engines over symmetric memory ("ce", the default where available) or as
# NCCL all-reduces ("nccl")
used = []
cases = [(None, sp, gr, False, "auto") for sp in (False, True) for gr in (False, True)]
cases += [(None, True, True, True, "auto")]
cases += [(torch.bfloat16, True, gr, fu, impl) for impl in ("auto", "nccl") for gr in (False, True) for fu in (False, True)]
for comm_dtype, sparse, graph, fused, impl in cases:
    p = run(sparse, graph, comm_dtype=comm_dtype, fused=fused, impl=impl)
    gathered = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(gathered, p)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        err = (p - ref).abs().max().item()
        mean_err = (p - ref).abs().mean().item()
        good = same and (err < 2e-4 if comm_dtype is None else (mean_err < 2e-5 and err < 5e-3))
        print("comm=%s/%s sparse=%d graph=%d fused=%d ranks_identical=%s max|p - single_process|=%.3e mean=%.3e %s" % (
            "fp32" if comm_dtype is None else "bf16", used[-1], sparse, graph, fused, same, err, mean_err, "ok" if good else "FAIL"),
            flush=True)
        ok = ok and good
dist.barrier()
if rank == 0:
    print("DP CHECK", "PASS" if ok else "FAIL", flush=True)
torch.cuda.synchronize()
os._exit(0 if ok else 1)
