"""CPU: host-side mirror of the reference interface — index/mask construction, config, state_dict
layout, error behaviour, and the no-CPU-fallback rule."""
import os

import numpy as np
import pytest
import torch

from meme_challenge_b200 import _lib
from meme_challenge_b200.model.meme_uniter import MemeUniter
from meme_challenge_b200.model.model import UniterConfig, UniterModel
from meme_challenge_b200.utils.utils import get_attention_mask, get_gather_index, pad_tensors
from oracle.make_golden import IMG_DIM, TINY


def test_gather_index_and_mask_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "index_mask.npz"))
    i = 0
    while "c%d_T" % i in g:
        tl, nb, T = g["c%d_txt_lens" % i].tolist(), g["c%d_num_bbs" % i].tolist(), int(g["c%d_T" % i])
        am = get_attention_mask(tl, nb)
        gi = get_gather_index(tl, nb, len(tl), T, am.shape[1])
        assert am.dtype == torch.float32 and gi.dtype == torch.int64
        assert np.array_equal(am.numpy(), g["c%d_attn_mask" % i])
        assert np.array_equal(gi.numpy(), g["c%d_gather_index" % i])
        i += 1


def test_gather_index_asserts_like_reference():
    with pytest.raises(AssertionError):
        get_gather_index([3, 4], [2], 2, 8, 10)


def test_pad_tensors():
    out = pad_tensors([torch.ones(2, 3), torch.ones(4, 3)])
    assert out.shape == (2, 4, 3) and out[0, 2:].abs().sum() == 0


def test_state_dict_layout_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny_meme_uniter.npz"))
    ref = {k[3:]: g[k].shape for k in g.files if k.startswith("sd.")}
    cfg = UniterConfig.from_dict(TINY)
    m = MemeUniter(UniterModel(cfg, IMG_DIM), cfg.hidden_size, 1)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())  # same names, same order
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref[k]) and v.dtype == torch.float32, k
    m.load_state_dict({k: torch.from_numpy(g["sd." + k]) for k in ref}, strict=True)


def test_base_model_parameter_count():
    cfg = UniterConfig(28996)
    m = MemeUniter(UniterModel(cfg, 2048), 768, 1)
    assert sum(p.numel() for p in m.parameters()) == 109899521
    assert len(m.state_dict()) == 212


def test_config_errors_match_reference():
    with pytest.raises(ValueError):
        UniterConfig(3.5)
    with pytest.raises(ValueError):
        UniterModel({"hidden_size": 8}, 2048)
    with pytest.raises(ValueError):
        UniterModel(UniterConfig(100, hidden_size=100, num_attention_heads=3), 64)


def test_from_pretrained_renames_gamma_beta(tmp_path, golden_dir):
    import json
    g = np.load(os.path.join(golden_dir, "tiny_meme_uniter.npz"))
    sd = {}
    for k in g.files:
        if k.startswith("sd.uniter_model."):
            name = k[len("sd.uniter_model."):]
            name = name.replace("LayerNorm.weight", "LayerNorm.gamma").replace("LayerNorm.bias", "LayerNorm.beta")
            sd[name] = torch.from_numpy(g[k])
    cfg_path = tmp_path / "cfg.json"
    cfg_path.write_text(json.dumps(TINY))
    m = UniterModel.from_pretrained(str(cfg_path), sd, img_dim=IMG_DIM)
    want = g["sd.uniter_model.encoder.layer.1.output.LayerNorm.weight"]
    assert np.array_equal(m.encoder.layer[1].output.LayerNorm.weight.detach().numpy(), want)
    bad = dict(sd)
    bad["pooler.dense.bias"] = torch.zeros(3)
    with pytest.raises(RuntimeError):
        UniterModel.from_pretrained(str(cfg_path), bad, img_dim=IMG_DIM)


def test_no_cpu_fallback(built_lib):
    """The product path refuses CPU tensors instead of silently computing elsewhere."""
    cfg = UniterConfig.from_dict(TINY)
    m = MemeUniter(UniterModel(cfg, IMG_DIM), cfg.hidden_size, 1).eval()
    from oracle.uniter_oracle import synth_batch
    b = synth_batch(2, 6, 4, img_dim=IMG_DIM, vocab=TINY["vocab_size"], min_txt=2, min_bb=2)
    with pytest.raises(_lib.B200UError):
        m(input_ids=b["input_ids"], position_ids=b["position_ids"], img_feat=b["img_feat"],
          img_pos_feat=b["img_pos_feat"], attention_mask=b["attn_mask"], gather_index=b["gather_index"],
          output_all_encoded_layers=False)


def test_bench_reference_arm_emits_one_contract_line():
    """`bench.py --impl reference` (the reference's own modules from oracle/_ref on the host cores, the oracle
    port when they are absent; no GPU needed) prints exactly ONE JSON line on stdout carrying the contract keys;
    everything else goes to stderr."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "memes/s" and d["higher_is_better"] is True
    assert d["metric"] == "UNITER-base fwd+bwd memes/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "memes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_synthetic_generators_match_the_oracle_generators():
    """The package's synthetic batches (bench.py / smoke / tools) are what the golden fixtures were made with."""
    from meme_challenge_b200.data.synthetic import synth_batch, synth_pretrain_batch
    from oracle import uniter_oracle as O
    for kw in (dict(B=4, T=12, R=10, seed=3, variable=True, img_dim=64, vocab=512, min_txt=2, min_bb=2),
               dict(B=3, T=16, R=20, seed=9)):
        a, b = O.synth_batch(**kw), synth_batch(**kw)
        for k in a:
            assert (torch.equal(a[k], b[k]) if torch.is_tensor(a[k]) else a[k] == b[k]), k
    a = O.synth_pretrain_batch(4, 12, 10, seed=77, img_dim=64, vocab=512, label_dim=11, min_txt=2, min_bb=2)
    b = synth_pretrain_batch(4, 12, 10, seed=77, img_dim=64, vocab=512, label_dim=11, min_txt=2, min_bb=2)
    for k in a:
        if torch.is_tensor(a[k]):
            assert torch.equal(a[k], b[k]), k


def test_metrics_against_reference_golden(golden_dir):
    """evaluate.standard_metrics_binary (rank-statistic AUROC, one-sort threshold search) reproduces the
    reference's data/metrics.py on the committed vectors, ties included."""
    import numpy as np
    from meme_challenge_b200 import evaluate as E
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    for i in range(int(g["n_cases"])):
        p, l = torch.from_numpy(g["c%d_probs" % i]), torch.from_numpy(g["c%d_labels" % i])
        m = E.standard_metrics_binary(p, l, add_optimal_acc=True)
        for k in ("accuracy", "recall", "precision", "F1", "aucroc", "optimal_threshold", "optimal_accuracy"):
            assert abs(m[k] - float(g["c%d_%s" % (i, k)])) <= 1e-6, (i, k, m[k], float(g["c%d_%s" % (i, k)]))
    # sklearn agrees on AUROC for a random vector with ties
    from sklearn.metrics import roc_auc_score
    torch.manual_seed(5)
    p = (torch.rand(500) * 20).round() / 20
    l = (p + 0.4 * torch.randn(500) > 0.5).long()
    assert abs(E.aucroc(p, l) - roc_auc_score(l.numpy(), p.numpy())) < 1e-12


def test_export_predictions_csv_format(tmp_path):
    from meme_challenge_b200 import evaluate as E
    path = E.export_predictions(str(tmp_path / "p.csv"), torch.tensor([7, 42]), torch.tensor([0.25, 0.75]),
                                labels=torch.tensor([0, 1]))
    assert open(path).read() == "id,proba,label,gt\n7,0.250000,0,0\n42,0.750000,1,1\n"


def test_collate_and_box_features_follow_the_reference():
    """data/meme_dataset.py:152-214 + data/dataset_template.py:100-113: zero-padded features, arange position
    ids, mask / gather index from the reference formulas (oracle restatement), 7-d boxes."""
    from meme_challenge_b200.data.pipeline import box7, collate_memes
    from oracle import uniter_oracle as O
    torch.manual_seed(1)
    bbox = torch.tensor([[10., 20., 110., 220.], [0., 0., 50., 40.]])
    p = box7(bbox, 200., 400., normalize=True)
    assert torch.allclose(p[0], torch.tensor([0.05, 0.05, 0.55, 0.55, 0.5, 0.5, 0.25]))
    nbb = [3, 5, 2]
    samples = [dict(img_feat=torch.randn(n, 8), img_pos_feat=torch.rand(n, 7), label=torch.tensor(i % 2),
                    data_id=torch.tensor(100 + i)) for i, n in enumerate(nbb)]
    input_ids = torch.randint(1, 50, (3, 6))
    tl = [4, 6, 2]
    for valid in (True, False):
        b = collate_memes(samples, input_ids, tl, pad_regions_are_valid=valid)
        assert b["img_feat"].shape == (3, 5, 8) and torch.equal(b["img_feat"][0, 3:], torch.zeros(2, 8))
        assert torch.equal(b["position_ids"], torch.arange(6).unsqueeze(0).repeat(3, 1))
        il = [5] * 3 if valid else nbb     # the reference reads the region count off the padded tensor
        am = O.get_attention_mask(tl, il)
        assert torch.equal(b["attn_mask"], am)
        assert torch.equal(b["gather_index"], O.get_gather_index(tl, il, 3, 6, am.shape[1]))
        assert torch.equal(b["ids"], torch.tensor([100, 101, 102]))


def test_reference_loader_runs_the_unmodified_reference():
    """oracle/ref_loader.py imports the reference modules (oracle/_ref when built, else the checkout) with the
    two shims; the reference MemeUniter runs a forward on the package's synthetic batch."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("no reference checkout and oracle/_ref not built")
    ns = ref_loader.load()
    from meme_challenge_b200.data.synthetic import synth_batch
    cfg = ns.model.UniterConfig.from_dict(TINY)
    torch.manual_seed(0)
    m = ns.meme_uniter.MemeUniter(ns.model.UniterModel(cfg, IMG_DIM), cfg.hidden_size, 1).eval()
    b = synth_batch(2, 6, 4, img_dim=IMG_DIM, vocab=TINY["vocab_size"], min_txt=2, min_bb=2)
    with torch.no_grad():
        out = m(input_ids=b["input_ids"], position_ids=b["position_ids"], img_feat=b["img_feat"],
                img_pos_feat=b["img_pos_feat"], attention_mask=b["attn_mask"], gather_index=b["gather_index"],
                output_all_encoded_layers=False)
    assert out.shape == (2, 1) and torch.isfinite(out).all()


def test_copy_engine_exchange_slice_plan_covers_every_bucket():
    """GradBuckets.slice_plan (train.py): the per-rank slices of every bucket tile it exactly, start 16-byte aligned,
    and the staging areas of different (bucket, source rank) pairs never overlap -- for any world size, with and
    without the dense embedding tail as an extra bucket."""
    from meme_challenge_b200.train import GradBuckets
    segments = [(0, 1000), (1000, 8088), (8088, 15176), (15176, 15176 + 7096)]   # embeddings + 3 "layers"
    n = segments[-1][1]
    for world in (2, 3, 4, 8, 16):
        for tail_lo in (None, 200):
            g16_lo, n16, plan, stage = GradBuckets.slice_plan(segments, world, n, tail_lo)
            assert g16_lo == (segments[1][0] if tail_lo is None else tail_lo) and n16 == n - g16_lo
            assert len(plan) == len(segments) - 1 + (tail_lo is not None)
            used = []
            for off, ln, sl, soff in plan:
                assert sl % 8 == 0 and sl * world >= ln and off % 8 == 0 and soff % 8 == 0
                covered = 0
                for r in range(world):
                    a = min(r * sl, ln)
                    b = min(a + sl, ln)
                    assert a % 8 == 0 or a == ln
                    covered += b - a
                assert covered == ln
                used.append((soff, soff + world * sl))
            used.sort()
            assert all(used[i][1] <= used[i + 1][0] for i in range(len(used) - 1)) and used[-1][1] == stage
            if tail_lo is not None:   # the tail bucket is the last plan entry and the first elements of the bf16 range
                assert plan[-1][0] == 0 and plan[-1][1] == segments[0][1] - tail_lo
