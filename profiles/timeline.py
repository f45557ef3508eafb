"""GPU timeline of steady-state hot-path steps (run on the GPU box; there is no nsys in the image).
    python profiles/timeline.py [steps] [--mlps] > gpurun_out/timeline.txt
torch.profiler (Kineto/CUPTI) over `steps` steps of bench.py's workload.  Prints per step: wall time, GPU-busy time
(sum of kernel + memcpy + memset durations), the idle share, the kernels by total time split into libb2a's and
PyTorch's, and the largest idle gaps with the kernel that follows each (which launch the GPU was waiting for)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

pipe = importlib.import_module("3danimals_b200.pipeline")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
steps = int(args[0]) if args else 5
dev = torch.device("cuda:0")
scene = pipe.SyntheticScene(grid_res=128, batch=16, image_res=256)
hp = pipe.HotPath(scene, dev, mlps="--mlps" in sys.argv)
g1, g2 = scene.upstream_grads()
d1, d2 = torch.from_numpy(g1).to(dev), torch.from_numpy(g2).to(dev)
for _ in range(5):
    hp.step(d1, d2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        hp.step(d1, d2)
    torch.cuda.synchronize()

evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
if not evs:
    print("no device events recorded")
    sys.exit(0)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print("steps %d: span %.1f us/step, GPU busy %.1f us/step (%.1f %%), %d device activities/step" %
      (steps, (t1 - t0) / steps, busy / steps, 100.0 * busy / (t1 - t0), len(evs) // steps))
agg = {}
for e in evs:
    a = agg.setdefault(e.name[:100], [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
ours = sum(v[1] for k, v in agg.items() if "anonymous" in k or "unnamed" in k or "b2a" in k)
print("libb2a kernels %.1f us/step, everything else %.1f us/step" % (ours / steps, (busy - ours) / steps))
print("%10s %6s  name" % ("us/step", "n/step"))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:70]:
    print("%10.1f %6.1f  %s" % (t / steps, c / steps, k))
gaps = []
for a, b in zip(evs[:-1], evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 0:
        gaps.append((g, a.name[:60], b.name[:60]))
gaps.sort(reverse=True)
print("idle gaps: total %.1f us/step; largest:" % (sum(g[0] for g in gaps) / steps))
for g, a, b in gaps[:40]:
    print("%8.1f us  after %-60s before %s" % (g, a, b))
