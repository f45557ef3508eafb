"""Build libb2a.so (the C-ABI CUDA library of this package) in-tree for sm_100a.

    python 3danimals_b200/build.py [--force] [--verbose]

Output: 3danimals_b200/csrc/_build/libb2a.so (git-ignored; travels to the GPU box with the gpurun snapshot).
Every file: -gencode arch=compute_100a,code=sm_100a  -lineinfo (ncu source pages).
Two arithmetic regimes, chosen per file:
  EXACT  (-fmad=false, IEEE div/sqrt): files whose results are compared BIT-EXACTLY with the oracle - triangle ids /
         barycentrics (raster.cu), extracted vertices and faces (marching_tets.cu), interpolated attributes
         (interpolate.cu), antialiased images (antialias.cu), quantile selection (estimate_bones.cu).  Every fp32 op is
         individually rounded; oracle/raster_ref.c is built with -ffp-contract=off to match.
  FAST   (-use_fast_math: FMA contraction, approximate div / sqrt / exp): files whose contract is the 1e-4 relative
         tolerance - fused g-buffer (gbuffer.cu), skinning (lbs.cu), vertex normals (normals.cu), light shading (shade.cu).  The IEEE division
         subroutine alone was 25 % of the g-buffer backward's instructions (profiles/).
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(SRC_DIR, "_build")
LIB = os.path.join(OUT_DIR, "libb2a.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

COMMON_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden"]
EXACT_FLAGS = ["-fmad=false"]
FAST_FLAGS = ["-use_fast_math"]
FAST_FILES = {"gbuffer.cu", "lbs.cu", "normals.cu", "shade.cu"}


def sources():
    return sorted(glob.glob(os.path.join(SRC_DIR, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(SRC_DIR, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h")) + [os.path.abspath(__file__)]


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(nvcc, src, verbose):
    obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
    flags = FAST_FLAGS if os.path.basename(src) in FAST_FILES else EXACT_FLAGS
    cmd = [nvcc] + COMMON_FLAGS + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return obj, res


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OUT_DIR, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile(nvcc, s, verbose), sources()))
    objs = []
    for obj, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building %s" % obj)
        objs.append(obj)
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lcudart"], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libb2a.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
