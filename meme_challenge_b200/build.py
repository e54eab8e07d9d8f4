"""In-tree build of libb200u.so (nvcc, sm_100a only).

`python -m meme_challenge_b200.build` compiles every csrc/*.cu to an object file (in parallel,
re-using objects whose sources are unchanged) and links meme_challenge_b200/libb200u.so. nvcc
cross-compiles without a GPU, so this runs in the CPU-only container; the .so travels to the GPU
box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200u.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _deps_digest(src):
    h = hashlib.sha1()
    for p in [src] + sorted(
        os.path.join(d, f)
        for d in (CSRC, os.path.join(HERE, "..", "include"))
        for f in os.listdir(d)
        if f.endswith((".cuh", ".h"))
    ):
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    name = os.path.splitext(os.path.basename(src))[0]
    obj = os.path.join(BUILD, name + ".o")
    stamp = obj + ".sha1"
    digest = _deps_digest(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as fh:
        fh.write(digest)
    with open(os.path.join(BUILD, name + ".ptxas.log"), "w") as fh:
        fh.write(r.stderr)
    return obj, r.stderr


def build(verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
