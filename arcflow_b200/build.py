"""Builds libarcflow_b200.so (sm_100a only) in-tree with nvcc.

The library is a plain C-ABI shared object (see include/arcflow_b200.h): no torch headers, cudart
linked statically, the driver entry point for TMA descriptors resolved at run time — so it loads on a
box without a GPU (the CPU test tier checks exactly that) and travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libarcflow_b200.so"
INCLUDE = PKG_DIR.parent / "include"

SOURCES = ["host.cu", "gemm.cu", "gemm_tn.cu", "attention.cu", "attention_bwd.cu", "elementwise.cu", "optim.cu", "vae.cu", "engine.cu", "c_api.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else "nvcc"


def _sources():
    return [CSRC / s for s in SOURCES if (CSRC / s).exists()]


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [INCLUDE / "arcflow_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


# what the last build() call in this process did: "compiled" (nvcc ran) or "reused" (sources + flags match the
# fingerprint of the library already in the tree, e.g. the prebuilt .so that travelled to the GPU box)
LAST_BUILD = None


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path."""
    global LAST_BUILD
    LIB_DIR.mkdir(exist_ok=True)
    stamp = LIB_DIR / "build.stamp"
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == fp:
        LAST_BUILD = "reused"
        return LIB_PATH
    objs = []
    procs = []
    for src in _sources():
        obj = LIB_DIR / (src.stem + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode != 0):
            print(f"--- {src.name}\n{out}", flush=True)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [_nvcc(), "-shared", "-cudart", "static", "-o", str(LIB_PATH), *map(str, objs)]
    subprocess.run(link, check=True)
    stamp.write_text(fp)
    LAST_BUILD = "compiled"
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(f"{LAST_BUILD} {path}")
