"""Build libb2a.so (the C-ABI CUDA library of this package) in-tree for sm_100a.

    python 3danimals_b200/build.py [--force] [--verbose]

Output: 3danimals_b200/csrc/_build/libb2a.so (git-ignored; travels to the GPU box with the gpurun snapshot).
Flags: -gencode arch=compute_100a,code=sm_100a  -lineinfo (ncu source pages)  -fmad=false (every fp32 op in the
decision-making code is individually rounded so index buffers are bit-reproducible against oracle/raster_ref.c,
which is built with -ffp-contract=off).
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(SRC_DIR, "_build")
LIB = os.path.join(OUT_DIR, "libb2a.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--threads", "0",
]


def sources():
    return sorted(glob.glob(os.path.join(SRC_DIR, "*.cu")))


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(SRC_DIR, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + sources() + ["-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libb2a.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
