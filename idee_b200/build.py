"""Build libidee_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m idee_b200.build [--force] [--verbose]

The library lands in ``idee_b200/lib/`` so that it travels with the source tree (it is git-ignored).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libidee_b200.so")
INCLUDE = os.path.join(ROOT, "include")

SOURCES = ["api.cu", "embed.cu", "swin_block.cu", "conv.cu", "conv_tc.cu", "conv16_umma.cu", "conv96_umma.cu", "conv96_wgrad_umma.cu", "lfq.cu", "lfq_k.cu", "losses.cu", "cnn_enc.cu"]
BASE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr"]
NVCC_FLAGS = BASE_FLAGS + ["-I", INCLUDE, "-I", CSRC]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; idee_b200 needs the CUDA toolkit to build its kernels")


def _fingerprint() -> str:
    """Hash of the sources and flags; independent of where the tree is checked out (the GPU box uses another path)."""
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(INCLUDE, "idee_b200.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode())
            h.update(fh.read())
    h.update(" ".join(BASE_FLAGS + SOURCES).encode())
    return h.hexdigest()


def _up_to_date(fp: str) -> bool:
    stamp = os.path.join(LIBDIR, "build.stamp")
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == fp


def up_to_date() -> bool:
    """True when the in-tree library was built from the current sources and flags."""
    return _up_to_date(_fingerprint())


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    fp = _fingerprint()
    if not force and _up_to_date(fp):
        return LIB
    import fcntl
    with open(os.path.join(LIBDIR, "build.lock"), "w") as lock:      # one builder at a time (ranks of a torchrun job)
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and _up_to_date(fp):
            return LIB
        return _build_locked(fp, verbose)


def _build_locked(fp: str, verbose: bool) -> str:
    stamp = os.path.join(LIBDIR, "build.stamp")
    nvcc = _nvcc()
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB + ".tmp"
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if os.path.exists(stamp):
        os.remove(stamp)
    os.replace(tmp, LIB)                                             # atomic: a concurrent loader never sees a partial file
    with open(stamp, "w") as fh:
        fh.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
