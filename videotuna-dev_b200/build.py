"""Build libb200vt.so (sm_100a only) in-tree with nvcc. Used by __graft_entry__.build() and by `python -m`.

The shared object lands next to this file so it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200vt.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libb200vt.so cannot be built")


def _sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if not os.path.isfile(os.path.join(root, f)):
                continue
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc: str, src: str) -> tuple[str, str]:
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link libb200vt.so. Returns the library path."""
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()
    srcs = _sources()
    objs, logs = [], []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for obj, log in ex.map(lambda s: _compile_one(nvcc, s), srcs):
            objs.append(obj)
            logs.append(log)
    with open(os.path.join(OBJ, "ptxas.log"), "w") as fh:
        fh.write("\n".join(logs))
    if verbose:
        for line in "\n".join(logs).splitlines():
            if "registers" in line or "spill" in line and "0 bytes spill stores" not in line:
                print(line)
    link = [nvcc, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC", "-cudart", "static"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
