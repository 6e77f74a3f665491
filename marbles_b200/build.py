"""Build the in-tree CUDA library (sm_100a only) with nvcc.

    python -m marbles_b200.build          # -> marbles_b200/libmarbles_b200.so

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmarbles_b200.so")
SOURCES = ["kernels.cu", "patch.cu", "amr.cu", "api.cu"]
# negative-result kernels (step variants 1-4 of round 1, the march step 8 of round 2): only with MBL_EXPERIMENTS=1
EXPERIMENT_SOURCES = [os.path.join("experiments", "fused.cu"), os.path.join("experiments", "march.cu")]
EXPERIMENT_HEADERS = [os.path.join("experiments", "experiments.cuh")]
HEADERS = ["lattice.cuh", "kernels.cuh", "patch.cuh", "bcic.cuh", "internal.cuh", os.path.join("..", "..", "include", "marbles_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS + EXPERIMENT_SOURCES + EXPERIMENT_HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    """One nvcc per source file, in parallel, then one link step."""
    if not force and not stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    experiments = os.environ.get("MBL_EXPERIMENTS") == "1"
    sources = SOURCES + (EXPERIMENT_SOURCES if experiments else [])
    flags = NVCC_FLAGS + (["-DMBL_EXPERIMENTS"] if experiments else []) + (extra or [])
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src: str):
        obj = os.path.join(objdir, os.path.basename(src).replace(".cu", ".o"))
        cmd = [nvcc()] + [f for f in flags if f != "-shared"] + ccbin + ["-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            print(" ".join(cmd))
        return obj, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=len(sources)) as ex:
        results = list(ex.map(compile_one, sources))
    for obj, res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed")
        if verbose:
            print(res.stdout + res.stderr)
    cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + ccbin + ["-o", LIB] + [o for o, _ in results]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc link failed")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True, extra=["-Xptxas", "-v"] if "-v" in sys.argv else None)
