"""Where does the HOST time of one hot-path step go?  (run on the GPU box)
    python profiles/host_breakdown.py [steps]
cProfile over `steps` steady-state steps of bench.py's workload; prints the top functions by cumulative time and the
GPU-busy fraction (sum of CUDA-event kernel time / wall time)."""
import cProfile
import importlib
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pipe = importlib.import_module("3danimals_b200.pipeline")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda:0")
scene = pipe.SyntheticScene(grid_res=128, batch=16, image_res=256)
hp = pipe.HotPath(scene, dev)
g1, g2 = scene.upstream_grads()
d1, d2 = torch.from_numpy(g1).to(dev), torch.from_numpy(g2).to(dev)
for _ in range(5):
    hp.step(d1, d2)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    hp.step(d1, d2)
torch.cuda.synchronize()
pr.disable()
wall = (time.perf_counter() - t0) / steps
s = io.StringIO()
ps = pstats.Stats(pr, stream=s).sort_stats("cumulative")
ps.print_stats(45)
print("wall per step (under cProfile): %.3f ms" % (wall * 1e3))
txt = s.getvalue()
print("\n".join(l[:170] for l in txt.splitlines()))
