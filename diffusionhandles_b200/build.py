"""Builds libdiffhandles_b200.so in-tree with nvcc for sm_100a (B200).  No torch dependency: the library
is a plain C-ABI shared object (include/dh_b200.h) loaded with ctypes.

    python -m diffusionhandles_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdiffhandles_b200.so")
STAMP = os.path.join(LIB_DIR, "build.stamp")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr"]
# bit-exact geometry / splat: never contract a*b+c into an FMA, IEEE division and sqrt (no fast math anywhere)
PER_FILE = {
    "dh_loss.cu": (["-DDH_LOSS_PHASE_TIMERS"] if os.environ.get("DH_LOSS_PHASE_TIMERS") else []) +
                  ([f"-DDH_LOSS_MAX_GROUPS={os.environ['DH_LOSS_MAX_GROUPS']}"] if os.environ.get("DH_LOSS_MAX_GROUPS") else []),
    "dh_geometry.cu": ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"],
    "dh_splat.cu": ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"],
    "dh_raster.cu": ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"],
    "dh_guided_step.cu": ["-fmad=false", "-prec-div=true"],      # the rounding sequence of the separate torch ops
}
SOURCES = ["dh_geometry.cu", "dh_splat.cu", "dh_masks.cu", "dh_warp.cu", "dh_loss.cu", "dh_loss_patch.cu", "dh_poisson.cu", "dh_raster.cu",
           "dh_guided_step.cu"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; libdiffhandles_b200 cannot be built")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/dh_b200.h", "../build.py"]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            with open(p, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    if not force and is_current():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH_FLAGS, *COMMON, *PER_FILE.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- {src} ---\n{out}", flush=True)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", LIB_PATH, *objs, "-lcudart"]
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
