"""Host time of one hot-path step by segment, WITHOUT a profiler (cProfile inflates Python-heavy code ~1.5x).
    python profiles/host_segments.py [steps]
perf_counter around the public API calls of HotPath.forward, around every autograd Function forward/backward of ops.py
and around every C-ABI call; no synchronisation is added, so each number is enqueue (host) time - except the one
device->host read in the extraction, which shows up inside `extract`."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pipe = importlib.import_module("3danimals_b200.pipeline")
ops = importlib.import_module("3danimals_b200.ops")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
acc = {}


def timed(label, fn):
    def w(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            d = acc.setdefault(label, [0.0, 0])
            d[0] += time.perf_counter() - t
            d[1] += 1
    return w


# public API segments
pipe.dmtet_mod.DMTet.extract = timed("api  dmtet.extract (incl. the size readback)", pipe.dmtet_mod.DMTet.extract)
pipe.mesh_mod.make_mesh = timed("api  make_mesh", pipe.mesh_mod.make_mesh)
pipe.skinning_mod.estimate_bones = timed("api  estimate_bones", pipe.skinning_mod.estimate_bones)
pipe.skinning_mod.skinning = timed("api  skinning", pipe.skinning_mod.skinning)
pipe.render_mod.render_mesh = timed("api  render_mesh", pipe.render_mod.render_mesh)
pipe.render_mod._sample_field = timed("       render_mesh: field sample (torch)", pipe.render_mod._sample_field)
pipe.FixedLight.shade = timed("       render_mesh: light.shade (torch)", pipe.FixedLight.shade)
# autograd nodes
for name in dir(ops):
    obj = getattr(ops, name)
    if isinstance(obj, type) and issubclass(obj, torch.autograd.Function) and obj is not torch.autograd.Function:
        obj.forward = staticmethod(timed("node %s.forward" % name, obj.forward))
        obj.backward = staticmethod(timed("node %s.backward" % name, obj.backward))
ops._call = timed("       C-ABI calls (binding + launches)", ops._call)
torch.empty = timed("       torch.empty", torch.empty)
torch.empty_like = timed("       torch.empty_like", torch.empty_like)
torch.zeros = timed("       torch.zeros", torch.zeros)
torch.zeros_like = timed("       torch.zeros_like", torch.zeros_like)
ops._f32 = timed("       _f32 checks", ops._f32)
ops._size = timed("       C-ABI size queries", ops._size)

dev = torch.device("cuda:0")
scene = pipe.SyntheticScene(grid_res=128, batch=16, image_res=256)
hp = pipe.HotPath(scene, dev)
g1, g2 = scene.upstream_grads()
d1, d2 = torch.from_numpy(g1).to(dev), torch.from_numpy(g2).to(dev)
fwd = timed("step forward (host)", hp.forward)


def step():
    hp.sdf.grad = None
    hp.angles.grad = None
    shaded, dino = fwd()
    t = time.perf_counter()
    torch.autograd.backward([shaded, dino], [d1, d2])
    d = acc.setdefault("step backward (host)", [0.0, 0])
    d[0] += time.perf_counter() - t
    d[1] += 1


torch.autograd.set_multithreading_enabled(False)      # what bench.py does by default
for _ in range(5):
    step()
torch.cuda.synchronize()
acc.clear()
t0 = time.perf_counter()
for _ in range(steps):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / steps
print("wall per step %.1f us (with the timers in place)" % (wall * 1e6))
print("%10s %8s  segment" % ("us/step", "n/step"))
for k, (t, n) in sorted(acc.items(), key=lambda x: (x[0].startswith(" "), -x[1][0])):
    print("%10.1f %8.1f  %s" % (t / steps * 1e6, n / steps, k))
