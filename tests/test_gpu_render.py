"""GPU parity of the drop-in API: render_mesh (reference model/render/render.py:228-337) and the whole hot path
(3danimals_b200.pipeline.HotPath) against the CPU oracle twin (oracle/torch_ref.render_mesh, oracle/pipeline_ref)."""
import numpy as np
import pytest
import torch

from conftest import pkg, rel_err
from oracle import pipeline_ref as P
from oracle import torch_ref as T

pytestmark = pytest.mark.gpu
TOL = 1e-4


def dev(x, cuda):
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


def _stage_inputs(grid_res=24, batch=2, image=64):
    pipe = pkg("pipeline")
    sc = pipe.SyntheticScene(grid_res=grid_res, batch=batch, image_res=image, sdf_noise=0.0)
    ref = P.forward(sc)            # CPU: extraction, bones, skinning -> posed vertices
    return pipe, sc, ref


@pytest.mark.parametrize("spp,modes,with_bg", [
    (1, ("shaded", "dino_pred"), False),
    (1, ("shaded", "shading", "kd", "geo_normal", "normal"), True),
    (2, ("shaded", "shading", "kd"), True),          # visualisation path: msaa, spp>1 (visualize_results.py:275-278)
    (1, ("shaded", "depth", "bogus"), False),
])
def test_render_mesh_matches_oracle(cuda, spp, modes, with_bg):
    """Identical posed vertices on both sides -> bit-exact triangle ids, images and all input gradients <= 1e-4."""
    mesh_mod, render_mod = pkg("render.mesh"), pkg("render.render")
    pipe, sc, ref = _stage_inputs()
    B, r = sc.batch, sc.image_res
    rng = np.random.RandomState(9)
    bg = rng.rand(B, r, r, 3).astype(np.float32) if with_bg else None
    posed = ref["posed"][:, 0].detach().clone().requires_grad_(True)
    prior = ref["verts"].detach()[None].clone().requires_grad_(True)
    faces = ref["faces"]
    w2c_t = torch.from_numpy(sc.w2c).requires_grad_(True)
    cam_t = torch.from_numpy(sc.campos).requires_grad_(True)
    mvp_t = torch.from_numpy(sc.mvp).requires_grad_(True)
    shader = P.analytic_shader(sc, B)

    def shade_fn(gb_tex, cam_normal, gbuf):
        d = shader(gb_tex, cam_normal, gbuf)
        if "depth" in modes:   # render.py:102-108
            gp = gbuf["gb_pos"]
            hom = torch.cat([gp, torch.ones_like(gp[..., :1])], -1)
            depth = torch.matmul(hom.view(B, -1, 4), w2c_t.transpose(-1, -2)).view(B, gp.shape[1], gp.shape[2], 4)[..., 2]
            mn, mx = depth.amin(dim=(1, 2), keepdim=True), depth.amax(dim=(1, 2), keepdim=True)
            d["depth"] = ((depth - mn) / (mx - mn)).unsqueeze(-1)
        return d

    known = [m for m in modes if m != "bogus"]
    v_nrm = T.auto_normals(posed, faces)
    out_ref = T.render_mesh(posed, v_nrm, faces, mvp_t, w2c_t, cam_t, shade_fn, (r, r), spp=spp,
                            background=torch.from_numpy(bg) if with_bg else None, render_modes=known, prior_v_pos=prior)
    gs = {k: rng.randn(*out_ref[k].shape).astype(np.float32) for k in known}
    sum((out_ref[k] * torch.from_numpy(gs[k])).sum() for k in known).backward()

    posed_d = dev(posed.detach().numpy(), cuda).requires_grad_(True)
    prior_d = dev(prior.detach().numpy(), cuda).requires_grad_(True)
    w2c_d, cam_d, mvp_d = (dev(x, cuda).requires_grad_(True) for x in (sc.w2c, sc.campos, sc.mvp))
    faces_d = dev(faces.numpy(), cuda)
    inst = mesh_mod.make_mesh(posed_d, faces_d[None], None, None, None)
    prior_mesh = mesh_mod.make_mesh(prior_d, faces_d[None], None, None, None)
    material = pipe.AnalyticField(dev(sc.w_kd, cuda), True)
    dino = pipe.AnalyticField(dev(sc.w_dino, cuda), False)
    light = pipe.FixedLight(dev(sc.light, cuda))
    outs = render_mod.render_mesh(None, inst, mvp_d, w2c_d, cam_d, material, light, (r, r), spp=spp, msaa=True,
                                  background=dev(bg, cuda) if with_bg else None, bsdf="diffuse", render_modes=list(modes),
                                  prior_mesh=prior_mesh, dino_net=dino)
    assert len(outs) == len(modes)
    total = 0
    for m, o in zip(modes, outs):
        if m == "bogus":
            assert o is None       # unknown keys give None (render.py:308-309)
            continue
        assert tuple(o.shape) == tuple(out_ref[m].shape), m
        assert rel_err(o.detach().cpu().numpy(), out_ref[m].detach().numpy()) < TOL, m
        total = total + (o * dev(gs[m], cuda)).sum()
    total.backward()
    for name, a, b in (("posed", posed_d, posed), ("prior", prior_d, prior), ("w2c", w2c_d, w2c_t), ("campos", cam_d, cam_t),
                       ("mvp", mvp_d, mvp_t)):
        if b.grad is None:
            assert a.grad is None or float(a.grad.abs().max()) == 0, name
            continue
        assert rel_err(a.grad.cpu().numpy(), b.grad.numpy()) < 3e-4, name


def test_render_mesh_asserts(cuda):
    mesh_mod, render_mod = pkg("render.mesh"), pkg("render.render")
    v = torch.rand(1, 5, 3, device=cuda)
    empty = mesh_mod.Mesh(v, torch.zeros(1, 0, 3, dtype=torch.long, device=cuda))
    with pytest.raises(AssertionError, match="empty training triangle mesh"):
        render_mod.render_mesh(None, empty, torch.eye(4, device=cuda)[None], torch.eye(4, device=cuda)[None], torch.zeros(1, 3, device=cuda),
                               None, None, (8, 8), render_modes=["shaded"], bsdf="diffuse")
    with pytest.raises(AssertionError):
        mesh_mod.make_mesh(v[0], torch.zeros(1, 1, 3, dtype=torch.long, device=cuda), None, None, None)   # unbatched verts


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: the product path fails loudly on host tensors."""
    ops, lib = pkg("ops"), pkg("_lib")
    with pytest.raises(lib.B2AError):
        ops.xfm_points(torch.zeros(1, 4, 3), torch.eye(4)[None])


@pytest.mark.parametrize("grid_res,batch,image", [(32, 2, 64), (48, 3, 128)])
def test_hot_path_matches_oracle(cuda, grid_res, batch, image):
    """Whole path (extraction -> bones -> LBS -> normals -> render -> backward).  Vertex positions now differ by fp32
    rounding between CPU and GPU (different exp / summation order in LBS), so a handful of edge pixels may pick the
    neighbouring triangle: allow <= 0.2 % of pixels to differ, everything else <= 1e-4.  Gradients: the antialias
    position gradient under white-noise upstream gradients is ill-conditioned - the ORACLE's own d_angles moves by
    0.3 % / 1.9 % when its posed vertices are perturbed by 1.2e-7 / 1e-6 relative (measured, DESIGN.md "Parity") - so
    the end-to-end gradient bar is 5e-2; the 1e-4-level gradient checks are the stage tests above, which feed both
    sides identical inputs."""
    pipe = pkg("pipeline")
    sc = pipe.SyntheticScene(grid_res=grid_res, batch=batch, image_res=image, sdf_noise=0.01)
    g1, g2 = sc.upstream_grads()
    d_sdf_ref, d_ang_ref, ref = P.step(sc, g1, g2)
    hp = pipe.HotPath(sc, cuda)
    d_sdf, d_ang = hp.step(dev(g1, cuda), dev(g2, cuda))
    prior = hp.last["prior"]
    assert np.array_equal(prior.t_pos_idx[0].cpu().numpy(), ref["faces"].numpy())               # bit-exact faces
    assert np.array_equal(prior.v_pos[0].detach().cpu().numpy(), ref["verts"].detach().numpy())
    assert [(b, list(d)) for b, d in hp.kinematic_chain] == [(b, list(d)) for b, d in ref["kinematic_chain"]]
    assert rel_err(hp.last["bones"].cpu().numpy(), ref["bones"].numpy()) < 1e-5
    assert rel_err(hp.last["inst"].v_pos.detach().cpu().numpy(), ref["posed"][:, 0].detach().numpy()) < TOL
    assert rel_err(hp.last["posed_bones"].detach().cpu().numpy(), ref["posed_bones"].detach().numpy()) < TOL
    shaded, dino = hp.forward()
    for o, k in ((shaded, "shaded"), (dino, "dino_pred")):
        a, b = o.detach().cpu().numpy(), ref[k].detach().numpy()
        bad = (np.abs(a - b).max(axis=1) > TOL * max(np.abs(b).max(), 1e-12))
        assert bad.mean() < 2e-3, (k, bad.mean())
    assert rel_err(d_ang.cpu().numpy(), d_ang_ref.numpy()) < 5e-2
    assert rel_err(d_sdf.cpu().numpy(), d_sdf_ref.numpy()) < 5e-2
