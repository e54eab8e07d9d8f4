"""CPU, world_size 2, gloo: the data-parallel gradient-bucket logic of train.GradBuckets — the
buckets tile the flat gradient buffer exactly once, are reduced in reverse layer order as the
backward proceeds, and leave every rank with the sum over ranks (the optimizer folds in 1/world)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _layout(n_layers=3):
    names, off, entries = [], 0, []
    sizes = {"uniter_model.embeddings.word_embeddings.weight": 800, "uniter_model.embeddings.LayerNorm.weight": 16,
             "uniter_model.img_embeddings.img_linear.weight": 256}
    for l in range(n_layers):
        sizes["uniter_model.encoder.layer.%d.attention.self.query.weight" % l] = 256
        sizes["uniter_model.encoder.layer.%d.output.dense.weight" % l] = 512
        sizes["uniter_model.encoder.layer.%d.output.LayerNorm.bias" % l] = 16
    sizes["uniter_model.pooler.dense.weight"] = 256
    sizes["linear.weight"] = 16
    for k, v in sizes.items():
        entries.append((k, off))
        off += v
    return entries, off


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from meme_challenge_b200.train import GradBuckets
    entries, n = _layout()
    torch.manual_seed(100 + rank)
    grad = torch.randn(n)
    mine = grad.clone()
    gb = GradBuckets(entries, grad, None, world)
    # segments tile [0, n) exactly once, embeddings first, last bucket takes pooler + head
    assert gb.segments[0][0] == 0 and gb.segments[-1][1] == n
    assert all(gb.segments[i][1] == gb.segments[i + 1][0] for i in range(len(gb.segments) - 1))
    assert len(gb.segments) == 4
    assert gb.segments[0][1] == 800 + 16 + 256
    # backward order: last layer first, embeddings last
    for layer in (2, 1, 0):
        gb.reduce_bucket(layer + 1)
    gb.reduce_bucket(0)
    order = gb.wait()
    assert order == [3, 2, 1, 0]
    others = [torch.zeros(n) for _ in range(world)]
    dist.all_gather(others, mine)
    want = sum(others)
    ok = torch.allclose(grad, want, rtol=1e-6, atol=1e-6)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_grad_buckets_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_cosine_schedule_matches_transformers():
    from meme_challenge_b200.train import cosine_with_warmup
    try:
        from transformers import get_cosine_schedule_with_warmup
    except Exception:  # pragma: no cover
        return
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sch = get_cosine_schedule_with_warmup(opt, num_warmup_steps=5, num_training_steps=40)
    for step in range(40):
        assert abs(opt.param_groups[0]["lr"] - cosine_with_warmup(step, 5, 40)) < 1e-6
        opt.step()
        sch.step()
