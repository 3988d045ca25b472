"""Build ``mulactseg_b200/lib/libmulactseg_b200.so`` (sm_100a only) with nvcc.

``python -m mulactseg_b200.build``.  Plain nvcc, no torch headers: the library
exposes the C ABI of ``include/mulactseg_b200.h`` and links cudart statically, so
it loads on a box without a GPU (the CPU test tier checks its exports) and
travels in-tree to the GPU box.  Objects are cached by source hash.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "csrc", "build")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.environ.get("MAS_LIB_OUT") or os.path.join(LIBDIR, "libmulactseg_b200.so")      # MAS_LIB_OUT: comparison builds

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-diag-suppress", "177"]
FLAGS += os.environ.get("MAS_NVCC_EXTRA", "").split()      # development: extra -D switches (comparison builds)


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path: str) -> str:
    h = hashlib.sha256()
    for dep in [path] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")] + \
            [os.path.join(ROOT, "include", "mulactseg_b200.h")]:
        with open(dep, "rb") as f:
            h.update(f.read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()[:16]


def _compile(src: str) -> str:
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, f"{os.path.splitext(src)[0]}.{_digest(path)}.o")
    if not os.path.exists(obj):
        for old in os.listdir(OBJ):
            if old.startswith(os.path.splitext(src)[0] + ".") and not os.environ.get("MAS_LIB_OUT"):
                os.remove(os.path.join(OBJ, old))
        cmd = [NVCC, *ARCH, *FLAGS, "-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(_compile, _sources()))
    stamp = os.path.join(OBJ, "link.stamp" if not os.environ.get("MAS_LIB_OUT") else "link.alt.stamp")
    want = " ".join(os.path.basename(o) for o in objs)
    have = open(stamp).read() if os.path.exists(stamp) else ""
    if want != have or not os.path.exists(LIB):
        cmd = [NVCC, *ARCH, "-shared", "-cudart", "static", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(want)
    if verbose:
        print(f"built {LIB} ({os.path.getsize(LIB) / 1e6:.1f} MB) from {len(objs)} objects")
    return LIB


if __name__ == "__main__":
    build()
    sys.exit(0)
