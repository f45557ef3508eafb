"""Where the HOST time of one M1a step goes, from torch.profiler's CPU + CUDA-runtime events (no Python-level timers):
    python profiles/host_profile.py [steps] > profiles/host_profile_rNN.txt
Totals per step of: CUDA runtime calls (cudaLaunchKernel, memset, event, stream waits ...), ATen ops by self CPU time, and the
wall time of the step with the profiler off."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

pipe = importlib.import_module("3danimals_b200.pipeline")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
scene = pipe.SyntheticScene(grid_res=128, batch=16, image_res=256)
hp = pipe.HotPath(scene, dev)
g1, g2 = scene.upstream_grads()
d1, d2 = torch.from_numpy(g1).to(dev), torch.from_numpy(g2).to(dev)
with torch.autograd.set_multithreading_enabled(False):
    for _ in range(5):
        hp.step(d1, d2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        hp.step(d1, d2)
    torch.cuda.synchronize()
    print("wall per step, profiler off: %.1f us" % ((time.perf_counter() - t0) / 100 * 1e6))
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            hp.step(d1, d2)
        torch.cuda.synchronize()
rt, ops_, gpu = {}, {}, 0.0
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        gpu += e.time_range.end - e.time_range.start
        continue
    name = e.name
    dur = e.time_range.end - e.time_range.start
    if name.startswith("cuda") or name.startswith("cu"):
        a = rt.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += dur
    else:
        a = ops_.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += e.self_cpu_time_total
print("GPU busy per step: %.1f us" % (gpu / steps))
print("CUDA runtime calls per step (host side): %.1f us total" % (sum(v[1] for v in rt.values()) / steps))
for k, (c, t) in sorted(rt.items(), key=lambda x: -x[1][1])[:10]:
    print("   %8.1f us x%-5.1f %s" % (t / steps, c / steps, k))
print("ATen / autograd ops by SELF CPU time per step: %.1f us total" % (sum(v[1] for v in ops_.values()) / steps))
for k, (c, t) in sorted(ops_.items(), key=lambda x: -x[1][1])[:25]:
    print("   %8.1f us x%-5.1f %s" % (t / steps, c / steps, k[:90]))
