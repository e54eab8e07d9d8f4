"""TEST INFRASTRUCTURE ONLY — recipe that makes the UNMODIFIED reference hot path runnable on the GPU box.

The reference is pure Python (no native code, no installable package), and /root/reference does not
exist on the GPU box. This script copies the files of the path that `BASELINE.json: north_star` names
byte for byte from the reference checkout into oracle/_ref/ (git-ignored, NOT gpurun-ignored, so it
travels with the snapshot exactly like a built .so) and records their SHA-256 in a manifest:

    model/{__init__,model,layer,meme_uniter,ot,pretrain}.py   the modules bench.py --impl reference runs
    utils/{__init__,optim_utils}.py                            get_optimizer: Adam + L2 decay groups
    config/uniter-{base,large}.json                            the two model configurations

Nothing is edited. The two import-time shims the modules need on this image (apex is
not installed -> FusedLayerNorm = torch.nn.LayerNorm; ot.trace's uint8 mask raises on torch >= 2) are
applied by oracle/ref_loader.py at import time, not to the files.

    python oracle/build_ref.py            # run in the build container; __graft_entry__.build() calls it
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B200U_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["model/__init__.py", "model/model.py", "model/layer.py", "model/meme_uniter.py", "model/ot.py",
         "model/pretrain.py", "utils/__init__.py", "utils/optim_utils.py", "config/uniter-base.json",
         "config/uniter-large.json"]


def build(verbose=False):
    """Copy the reference files if the checkout is present. Returns the manifest path or None."""
    if not os.path.isdir(REF):
        return None
    manifest = {"source": REF, "files": {}}
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest["files"][rel] = hashlib.sha256(fh.read()).hexdigest()
        if verbose:
            sys.stderr.write("copied %s\n" % rel)
    path = os.path.join(DST, "MANIFEST.json")
    with open(path, "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    return path


if __name__ == "__main__":
    p = build(verbose=True)
    print(p if p else "reference checkout %s not found: nothing copied" % REF)
