"""Host timing of the OBJ export at the BASELINE configs[4] mesh size (85 k vertices, 170 k faces, 200 k
texcoords - DESIGN.md §6): the reference's per-line loop (restated in oracle/obj_text.py; the reference's own function when the
tree is present) against `3danimals_b200.render.obj.obj_text` (libb2a.so, b2a_obj_format) at 1 thread and at all threads.
Best of 5, perf_counter, byte equality asserted.   python profiles/obj_export.py > profiles/obj_export_r1.txt"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import obj_text as oracle_obj  # noqa: E402

obj = importlib.import_module("3danimals_b200.render.obj")


def best(fn, n=9):
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t)
    return min(ts), r


def main():
    rng = np.random.default_rng(0)
    V, Vt, F = 85_000, 200_000, 170_000
    v_pos = (rng.standard_normal((V, 3)) * 0.5).astype(np.float32)
    v_nrm = rng.standard_normal((V, 3)).astype(np.float32)
    v_nrm /= np.linalg.norm(v_nrm, axis=-1, keepdims=True)
    v_tex = rng.random((Vt, 2)).astype(np.float32)
    t_pos = rng.integers(0, V, (F, 3))
    t_tex = rng.integers(0, Vt, (F, 3))
    t_ref, want = best(lambda: oracle_obj.obj_text(v_pos, t_pos, v_nrm, t_pos, v_tex, t_tex), n=2)
    print("host cores: %d" % os.cpu_count())
    print("mesh: V=%d Vt=%d F=%d -> %.1f MB of text, %d lines" % (V, Vt, F, len(want) / 1e6, want.count(b"\n")))
    print("reference per-line loop (oracle/obj_text.py, 1 thread):  %8.1f ms" % (t_ref * 1e3))
    for threads in (1, 2, 4, 8, 0):
        t, got = best(lambda: obj.obj_text(v_pos, t_pos, v_nrm, t_pos, v_tex, t_tex, threads=threads))
        assert got.tobytes() == want
        label = "all" if threads == 0 else str(threads)
        print("b2a_obj_format, threads=%-3s  %8.2f ms   %6.1f MB/s of text   %5.0fx" % (label, t * 1e3, len(want) / t / 1e6, t_ref / t))


if __name__ == "__main__":
    main()
