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
    """`bench.py --impl reference` (the reference's CPU path = the oracle port, no GPU needed) prints exactly
    ONE JSON line on stdout carrying the contract keys; everything else goes to stderr."""
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
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "memes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
