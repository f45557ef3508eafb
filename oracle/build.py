"""Build the C oracle (TEST INFRASTRUCTURE ONLY) -> oracle/_build/liboracle.so.

-ffp-contract=off keeps every fp32 op individually rounded, matching the CUDA product's -fmad=false, so the
triangle-id buffer is bit-reproducible across CPU and GPU.  No -march flags: the .so is built in the build
container and travels to the GPU box (different host CPU).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")
SRC = os.path.join(HERE, "raster_ref.c")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-std=gnu11", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden",
           "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
