"""Builds compseed_b200/_lib/libcompseed_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
# experiment builds: CS_DEFS="-DCS_LIST_SMEM=12 -DCS_SEED_MINBLOCKS=4" CS_TAG=e12b4 python -m compseed_b200.build
TAG = os.environ.get("CS_TAG", "")
EXTRA_DEFS = os.environ.get("CS_DEFS", "").split()
LIB = os.path.join(LIB_DIR, f"libcompseed_b200{('_' + TAG) if TAG else ''}.so")
SOURCES = ["cs_kernels.cu", "cs_api.cu", "cs_index_build.cu", "cs_verify.cu", "cs_multi.cu", "cs_chain.cu", "cs_bsw.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr", "-rdc=false"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "compseed_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", (("_" + TAG) if TAG else "") + ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *EXTRA_DEFS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-lcudart"])   # (arch also at link time: no stub for nvcc's default architecture)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
