"""BASELINE configs[4] shape (visualize rotation / texture finetune): B=1, DMTet res 256 (17 M grid vertices, 100 M tets),
512x512 at spp 4 (2048^2 internal), render modes of visualize_results.py.  Not a bench line - a per-call breakdown of
where the time goes at that size (CUDA events around every C-ABI call + wall time of the phases).
    python profiles/c4_step.py > gpurun_out/c4_step.txt
Phases timed: (1) extraction + normals + bones + skinning of the res-256 shape (once per rotation sequence);
(2) one rotation frame: render ['shaded','shading','kd'] at 512^2 x spp 4, forward only (visualize_results.py:353-396);
(3) one texture-finetune iteration: render ['shaded'] + backward to the texture stand-in's input (fixed geometry)."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pipe = importlib.import_module("3danimals_b200.pipeline")
ops = importlib.import_module("3danimals_b200.ops")
syn = importlib.import_module("3danimals_b200.synthetic")
mesh_mod = importlib.import_module("3danimals_b200.render.mesh")
render_mod = importlib.import_module("3danimals_b200.render.render")
sk = importlib.import_module("3danimals_b200.geometry.skinning")
dm = importlib.import_module("3danimals_b200.geometry.dmtet")

dev = torch.device("cuda:0")
RES, IMG, SPP = 256, 512, 4
v, t = syn.kuhn_tet_grid_torch(RES, dev)
v = (v * 7.0).contiguous()
sdf = torch.from_numpy(syn.sdf_horse(v.cpu().numpy(), sigma=0.0)).to(dev)[:, None].contiguous()
mt = dm.DMTet()
t0 = time.perf_counter()
grid = mt.grid_for(t, v.shape[0])
torch.cuda.synchronize()
print("static grid tables (once per grid): %.2f s; Vg %d  T %d  E %d" % (time.perf_counter() - t0, grid.Vg, grid.T, grid.E))
del t
mvp, w2c, campos = (torch.from_numpy(x).to(dev) for x in syn.cameras(1, seed=3))
rng = np.random.RandomState(0)
material = pipe.AnalyticField(torch.from_numpy((rng.randn(3, 3) * 1.5).astype(np.float32)).to(dev), True)
light = pipe.FixedLight(torch.tensor([0.3, 0.5, 0.8, 0.4, 0.6], device=dev))
angles = torch.from_numpy(rng.uniform(-0.3, 0.3, size=(1, 1, 20, 3)).astype(np.float32)).to(dev)


def geometry():
    verts, faces, uv_idx, faces32 = mt.extract(v, sdf, grid)
    prior = mesh_mod.make_mesh(verts[None], faces[None], None, uv_idx[None], None, faces_i32=faces32)
    bones, chain, aux = sk.estimate_bones(prior.v_pos[:, None].detach(), 8, n_legs=4, n_leg_bones=3, body_bones_mode="z_minmax_y+",
                                          compute_kinematic_chain=True)
    posed, _ = sk.skinning(prior.v_pos[:, None], bones, chain, angles, output_posed_bones=True, temperature=0.05)
    inst = mesh_mod.make_mesh(posed[:, 0], prior.t_pos_idx, None, prior.t_tex_idx, None, faces_i32=prior.tri_i32())
    return prior, inst


def frame(prior, inst, modes):
    return render_mod.render_mesh(None, inst, mvp, w2c, campos, material, light, (IMG, IMG), spp=SPP, num_layers=1, msaa=True,
                                  background=None, bsdf="diffuse", render_modes=list(modes), prior_mesh=prior)


def timed(label, fn, reps=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    ops.stats.reset()
    ops.stats.timing = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    ops.stats.timing = False
    durs = ops.stats.durations_ms()
    ours = sum(sum(x) for x in durs.values()) / reps
    print("\n== %s: %.3f ms wall, %.3f ms device span, libb2a calls %.3f ms" % (label, wall * 1e3, e0.elapsed_time(e1) / reps, ours))
    for (n, tag), x in sorted(durs.items(), key=lambda kv: -sum(kv[1])):
        print("   %8.1f us x%-3d %s %s" % (sum(x) / len(x) * 1e3, len(x) // reps, n, tag))
    return out


with torch.no_grad():
    prior, inst = timed("geometry: extraction (res 256) + normals + bones + LBS + normals", geometry, reps=3)
    print("   mesh: V %d  F %d" % (inst.v_pos.shape[1], inst.t_pos_idx.shape[1]))
    timed("rotation frame: ['shaded','shading','kd'] 512^2 x spp4, forward", lambda: frame(prior, inst, ("shaded", "shading", "kd")))

tex_in = None


def finetune():
    outs = frame(prior, inst, ("shaded",))
    loss = (outs[0] ** 2).mean()
    g = torch.autograd.grad(loss, [material_param])
    return g


# texture finetune: gradient w.r.t. the texture field's weights only (geometry fixed)
material_param = material.weight.clone().requires_grad_(True)


class _TexField(torch.nn.Module):
    bsdf = None
    dense_only = True

    def sample(self, x, feat=None):
        y = torch.sigmoid(x @ material_param)
        return torch.cat([y, y, y], -1)


material = _TexField()
timed("texture-finetune iteration: ['shaded'] 512^2 x spp4, fwd+bwd to the texture weights", finetune)

# ---- the same two loops replayed from CUDA graphs (3danimals_b200/graphs.py) ------------------------------------------
graphs = importlib.import_module("3danimals_b200.graphs")
material = pipe.AnalyticField(material_param.detach().clone(), True)
cap = graphs.captured_render(inst, prior, material, light, (IMG, IMG), (mvp, w2c, campos), spp=SPP, render_modes=("shaded", "shading", "kd"))


def replay_timed(label, fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("\n== %s: %.3f ms wall, %.3f ms device span per iteration" % (label, (time.perf_counter() - t0) / reps * 1e3, e0.elapsed_time(e1) / reps))


replay_timed("rotation frame, CUDA-graph replay (new camera copied in every frame)", lambda: cap(mvp, w2c, campos))
material = _TexField()
target = torch.rand(1, 4, IMG, IMG, device=dev)


def finetune_it(tgt):
    outs = frame(prior, inst, ("shaded",))
    loss = ((outs[0] - tgt) ** 2).mean()
    g, = torch.autograd.grad(loss, [material_param])
    return loss.detach(), g


step = graphs.CapturedStep(finetune_it, [target])
replay_timed("texture-finetune iteration (fwd+bwd to the texture weights), CUDA-graph replay", lambda: step(target))

# ---- which kernels make up a rotation frame at this size (torch.profiler, eager) ---------------------------------------
from torch.profiler import ProfilerActivity, profile  # noqa: E402

material = pipe.AnalyticField(material_param.detach().clone(), True)
with torch.no_grad():
    frame(prior, inst, ("shaded", "shading", "kd"))
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            frame(prior, inst, ("shaded", "shading", "kd"))
        torch.cuda.synchronize()
agg = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        a = agg.setdefault(e.name[:110], [0, 0.0])
        a[0] += 1
        a[1] += e.time_range.end - e.time_range.start
print("\n== rotation frame: device activities per frame (eager)")
for k, (c, tt) in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
    print("   %8.1f us x%-4.1f %s" % (tt / 3, c / 3, k))
