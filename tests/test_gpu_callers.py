"""The drop-in under the reference's REAL callers (north_star: "drops in under model/models and model/predictors unchanged").

The reference's own files - `model/predictors/BasePredictorBase.py`, `model/predictors/InstancePredictorBase.py`,
`model/models/AnimalModel.py`, `model/models/Fauna.py` - are imported UNMODIFIED from the staged tree (`oracle/stage_ref.py` ->
git-ignored `oracle/_ref/`, which travels to the GPU box with the snapshot) on top of `3danimals_b200.overlay`, executed on
the B200, and their outputs / gradients are compared with the CPU oracle evaluated on the same parameters
(`oracle/torch_ref.py`: restated reference geometry, pinned by the reference's goldens, + the C restatement of the rasterizer ops).

What is NOT the reference's code in these tests: the stubs of absent third-party packages (import level), the ViT encoder
(needs torch.hub: features are synthesised), and `nn.Module.__init__` instead of the predictor / model constructors that
would build that encoder - the methods under test (`BasePredictorBase.forward`, `forward_articulation`, `get_bones`,
`apply_articulation_constraints`, `AnimalModel.render`, `FaunaModel.get_random_view_mask`) run as shipped.
"""
import copy
import types

import numpy as np
import pytest
import torch

from conftest import pkg, rel_err
from oracle import geometry_np as gnp
from oracle import ref_callers as RC
from oracle import stage_ref
from oracle import torch_ref as T

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not RC.available(), reason="no staged reference tree (python oracle/stage_ref.py)")]

GRID_RES, IMG, BATCH, SCALE = 32, 64, 3, 7.0


@pytest.fixture(scope="module")
def ref(cuda, tmp_path_factory):
    import os
    ns = RC.load()
    cwd = os.getcwd()
    os.chdir(tmp_path_factory.mktemp("tets"))          # DMTetGeometry reads data/tets/<res>_tets.npz relative to the cwd (dmtet.py:223)
    os.environ["B2A_SYNTHETIC_TETS"] = "1"             # no downloaded grid offline: explicit opt-in to the synthetic Kuhn grid
    try:
        yield ns
    finally:
        os.environ.pop("B2A_SYNTHETIC_TETS", None)
        os.chdir(cwd)
        RC.unload()


def test_staged_tree_is_the_reference(ref):
    if ref.root == stage_ref.DEST:
        assert stage_ref.verify() == []
    assert ref.IPB.estimate_bones.__module__ == "3danimals_b200.geometry.skinning"
    assert ref.IPB.skinning.__module__ == "3danimals_b200.geometry.skinning"
    assert ref.BPB.DMTetGeometry.__module__ == "3danimals_b200.geometry.dmtet"
    assert ref.AM.render.__name__ == "3danimals_b200.render.render" and ref.AM.dr.__name__ == "3danimals_b200.nvdiffrast_shim.torch"
    assert ref.IPB.InstancePredictorBase.__module__ == "model.predictors.InstancePredictorBase"
    assert ref.IPB.__file__.startswith(ref.root) and ref.AM.__file__.startswith(ref.root)


def _base_predictor(ref, cuda, hidden=64):
    """The reference's BasePredictorBase, constructed by its own constructor (BasePredictorBase.py:44-66): DMTetGeometry
    (the drop-in, via the overlay) + the DINO field CoordMLP."""
    B = ref.BPB
    cfg = B.BasePredictorConfig(
        cfg_shape=B.DMTetConfig(grid_res=GRID_RES, spatial_scale=SCALE, num_layers=5, hidden_size=hidden, embedder_freq=8, embed_concat_pts=True,
                                init_sdf="ellipsoid", jitter_grid=0.0, symmetrize=True, grid_res_coarse_iter_range=[0, 0], grid_res_coarse=GRID_RES),
        cfg_dino=B.netDINOConfig(feature_dim=16, num_layers=5, hidden_size=hidden, activation="sigmoid", embedder_freq=8, symmetrize=True))
    torch.manual_seed(0)
    base = B.BasePredictorBase(cfg).to(cuda)
    # the ellipsoid init has no legs: bend the SDF network's output layer a little so that the shape is not a pure ellipsoid
    with torch.no_grad():
        for p in base.netShape.mlp.parameters():
            p.mul_(1.5)
    return base


def _instance_predictor(ref, cuda, n_body=8, n_legs=4, n_leg_bones=3, mode="z_minmax_y+", feat_dim=32):
    """InstancePredictorBase without its constructor (it builds a ViT through torch.hub); every attribute the articulation
    methods read is what the constructor would have set (InstancePredictorBase.py:199-233) for the MagicPony horse config."""
    I = ref.IPB
    pred = RC.bare(I.InstancePredictorBase)
    pred.cfg_articulation = I.ArticulationConfig(
        architecture="attention", num_layers=2, hidden_size=feat_dim, embedder_freq=8, bone_feature_mode="sample+global", num_body_bones=n_body,
        body_bones_mode=mode, num_legs=n_legs, num_leg_bones=n_leg_bones, attach_legs_to_body_iter_range=[0, 1 << 30], static_root_bones=False,
        skinning_temperature=0.05, max_arti_angle=60.0, constrain_legs=n_legs > 0, output_multiplier=0.1)
    pred.cfg_pose = I.PoseConfig(rot_rep="quadlookat")
    pred.spatial_scale = SCALE
    pred.enable_articulation = True
    pred.num_bones = n_body + n_legs * n_leg_bones
    torch.manual_seed(1)
    pred.netArticulation = ref.networks.ArticulationNetwork("attention", feat_dim * 2, 1 + 2 + 3 * 2, 2, feat_dim, n_harmonic_functions=8,
                                                            embedder_scalar=np.pi).to(cuda)
    pred.kinematic_tree_epoch = -1
    return pred


def _cameras(cuda, n, seed=3):
    syn = pkg("synthetic")
    return tuple(torch.from_numpy(x).to(cuda) for x in syn.cameras(n, seed=seed))


def _oracle_shader(texture, dino, light, feat):
    """The per-pixel shader of render.shade (render.py:50-94) on the host with copies of the SAME modules: texture field
    (feat = im_features), DINO field, DirectionalLight."""
    if light is not None:
        light.__dict__.pop("light_params", None)        # forward() stashes its (non-leaf) output on the module: not deep-copyable
    texture, dino, light = (copy.deepcopy(m).cpu() if m is not None else None for m in (texture, dino, light))
    feat = feat.detach().cpu() if feat is not None else None

    def shade(gb_tex, cam_normal, gbuf):
        all_tex = texture.sample(gb_tex, feat=feat) if texture is not None else torch.ones(*gb_tex.shape[:-1], 9)
        kd = all_tex[..., :3]
        out = {"kd": kd}
        if light is not None:
            lp = light.forward(feat)
            out["shaded"], out["shading"] = T.directional_shade(lp.reshape(-1, 5), kd, cam_normal)
        else:
            out["shaded"] = kd
        if dino is not None:
            out["dino_pred"] = dino.sample(gb_tex)
        return out

    return shade, (texture, dino, light)


def test_magicpony_chain_through_reference_callers(ref, cuda):
    """BasePredictorBase.forward -> InstancePredictorBase.forward_articulation -> AnimalModel.render, all the reference's own
    methods, on the B200 over libb2a.so; images and gradients (SDF network, articulation network, texture / DINO / light
    networks) against the CPU oracle driven by the same parameters."""
    base = _base_predictor(ref, cuda)
    pred = _instance_predictor(ref, cuda)
    I, A = ref.IPB, ref.AM
    torch.manual_seed(2)
    feat_dim = 32
    mm = torch.tensor([[0., 1.]] * 9, device=cuda)
    texture = ref.networks.CoordMLP(3, 9, 4, nf=64, activation="sigmoid", min_max=mm, n_harmonic_functions=10,
                                    embedder_scalar=2 * np.pi / SCALE * 0.9, extra_feat_dim=feat_dim, symmetrize=True).to(cuda)
    light_mod = __import__("importlib").import_module("model.render.light")
    light = light_mod.DirectionalLight(feat_dim, 3, 32, intensity_min_max=torch.tensor([[0.0, 1.0], [0.5, 1.0]])).to(cuda)
    feat = torch.randn(BATCH, feat_dim, device=cuda)
    patch = torch.randn(BATCH, feat_dim, 8, 8, device=cuda)
    mvp, w2c, campos = _cameras(cuda, BATCH)

    # ---- the reference's chain, unmodified --------------------------------------------------------------------
    prior_shape, net_dino = base.forward(total_iter=0, is_training=False)                       # R1
    shape, arti, aux = pred.forward_articulation(prior_shape, feat, patch, mvp, w2c, BATCH, 1, epoch=0, total_iter=10)   # R4 R5 R3
    model = RC.bare(A.AnimalModel)
    model.cfg_render = A.RenderConfig(spatial_scale=SCALE, background_mode="none", renderer_spp=1)
    model.glctx = None
    shaded, dino_pred = model.render(["shaded", "dino_pred"], shape, texture, mvp, w2c, campos, (IMG, IMG), im_features=feat, light=light,
                                     prior_shape=prior_shape, dino_net=net_dino)              # R6-R10
    assert tuple(shaded.shape) == (BATCH, 4, IMG, IMG) and tuple(dino_pred.shape) == (BATCH, 16, IMG, IMG)
    assert type(model.glctx).__module__ == "3danimals_b200.nvdiffrast_shim.torch"
    rng = np.random.RandomState(7)
    # smooth upstream gradients (low-pass noise): a well-conditioned end-to-end gradient check, unlike white noise through the
    # antialias position gradient
    def smooth(c):
        g = torch.from_numpy(rng.randn(BATCH, c, IMG // 8, IMG // 8).astype(np.float32))
        return torch.nn.functional.interpolate(g, size=(IMG, IMG), mode="bilinear", align_corners=False) * 1e-2
    g_sh, g_di = smooth(4), smooth(16)
    nets = dict(sdf=base.netShape.mlp, arti=pred.netArticulation, tex=texture, dino=net_dino, light=light.mlp)
    for m in nets.values():
        m.zero_grad()
    torch.autograd.backward([shaded, dino_pred], [g_sh.to(cuda), g_di.to(cuda)])

    # ---- oracle on the same parameters ------------------------------------------------------------------------
    sdf_net = copy.deepcopy(base.netShape.mlp).cpu()
    arti_net = copy.deepcopy(pred.netArticulation).cpu()
    for m in (sdf_net, arti_net):
        m.zero_grad()
    grid_v = base.netShape.verts.cpu()
    tets = base.netShape.indices.cpu()
    xs, ys, zs = grid_v.unbind(-1)
    sym = torch.stack([xs.abs(), ys, zs], -1)
    sdf = sdf_net(sym) + (SCALE * 0.15 - torch.stack([sym[:, 0], sym[:, 1], sym[:, 2] / 2], -1).norm(dim=-1, keepdim=True))   # dmtet.py:228-254
    verts, faces, uv_idx = T.marching_tets(grid_v, sdf, tets)
    assert np.array_equal(prior_shape.t_pos_idx[0].cpu().numpy(), faces.numpy())                                   # bit-exact topology
    # (the SDF network runs on the tensor-core field-MLP path: ~5e-6 on the SDF values, amplified by the zero-crossing interpolation)
    assert rel_err(prior_shape.v_pos[0].detach().cpu().numpy(), verts.detach().numpy()) < 1e-4
    bones, chain, baux = gnp.estimate_bones(verts.detach().numpy()[None, None], 8, n_legs=4, n_leg_bones=3, body_bones_mode="z_minmax_y+")
    assert [(b, list(d)) for b, d in pred.kinematic_tree] == [(b, list(d)) for b, d in chain]
    bones_t = torch.from_numpy(bones)
    assert rel_err(aux["bones_pred"].cpu().numpy() if "bones_pred" in aux else bones, bones) < 1e-5
    # get_bones' feature assembly (InstancePredictorBase.py:337-383) restated on the host with the oracle's bones
    mvp_c, w2c_c = mvp.cpu(), w2c.cpu()
    K = bones_t.shape[2]
    bp = bones_t.repeat(BATCH, 1, 1, 1, 1).view(BATCH, K, 2, 3)
    mid = bp.mean(2)
    mid4 = torch.cat([mid, torch.ones_like(mid[..., :1])], -1) @ mvp_c.transpose(-1, -2)
    mid2d = mid4[..., :2] / mid4[..., 3:4]
    cam4 = torch.cat([bp, torch.ones_like(bp[..., :1])], -1) @ w2c_c[:, None].transpose(-1, -2)
    cam3 = cam4[..., :3] / cam4[..., 3:4] + torch.tensor([0, 0, 10.0]).view(1, 1, 1, 3)
    pos3d = cam3.view(BATCH, K, 6) / SCALE * 2
    idx_in = ((torch.arange(K)[None, :, None] + 0.5) / K * 2 - 1).repeat(BATCH, 1, 1)
    pos_in = torch.cat([mid2d, pos3d, idx_in], -1)
    local = torch.nn.functional.grid_sample(patch.cpu(), mid2d.view(BATCH, 1, -1, 2), mode="bilinear", align_corners=False).squeeze(-2).permute(0, 2, 1)
    bones_feat = torch.cat([feat.cpu()[:, None].repeat(1, K, 1), local], -1)
    ang = arti_net(bones_feat, pos_in).view(BATCH, 1, K, 3)
    # apply_articulation_constraints (InstancePredictorBase.py:435-511) for this config: multiplier, tanh, leg twist / side-bend limits
    ang = (ang * 0.1).tanh()
    legs = 8 + np.arange(12)
    m = torch.zeros_like(ang); m[:, :, legs, 2] = 1
    ang = m * (ang * 0.3) + (1 - m) * ang
    m = torch.zeros_like(ang); m[:, :, legs, 1] = 1
    ang = m * (ang * 0.3) + (1 - m) * ang
    ang = ang * 60.0 / 180 * np.pi
    assert rel_err(arti.detach().cpu().numpy(), ang.detach().numpy()) < 1e-4
    posed, saux = T.skinning(verts[None, None], bones_t, chain, ang, temperature=0.05)
    assert rel_err(shape.v_pos.detach().cpu().numpy(), posed[:, 0].detach().numpy()) < 1e-4
    assert rel_err(aux["posed_bones"].detach().cpu().numpy(), saux["posed_bones"].detach().numpy()) < 1e-4
    v_nrm = T.auto_normals(posed[:, 0], faces)
    shade, (tex_c, dino_c, light_c) = _oracle_shader(texture, net_dino, light, feat)
    for m_ in (tex_c, dino_c, light_c):
        m_.zero_grad()
    out = T.render_mesh(posed[:, 0], v_nrm, faces, mvp_c, w2c_c, campos.cpu(), shade, (IMG, IMG), spp=1, background=torch.zeros(BATCH, IMG, IMG, 3),
                        render_modes=("shaded", "dino_pred"), prior_v_pos=verts[None])
    for got, key in ((shaded, "shaded"), (dino_pred, "dino_pred")):
        a, b = got.detach().cpu().numpy(), out[key].detach().numpy()
        bad = np.abs(a - b).max(axis=1) > 1e-4 * max(np.abs(b).max(), 1e-12)
        assert bad.mean() < 2e-3, (key, float(bad.mean()))
    torch.autograd.backward([out["shaded"], out["dino_pred"]], [g_sh, g_di])
    ref_nets = dict(sdf=sdf_net, arti=arti_net, tex=tex_c, dino=dino_c, light=light_c.mlp)
    worst = {}
    for name, net in nets.items():
        for (pn, p), (_, q) in zip(net.named_parameters(), ref_nets[name].named_parameters()):
            assert p.grad is not None, (name, pn)
            a, b = p.grad.cpu().double(), q.grad.double()
            worst[name] = max(worst.get(name, 0.0), float((a - b).norm() / b.norm().clamp_min(1e-30)))
    # Relative L2 error per parameter tensor (a ReLU pre-activation within fp32 rounding of zero may take the other branch on the
    # two sides, which moves single entries of the earlier layers' gradients - see test_sparse_field_evaluation_matches_dense).
    # Field / light networks see the same pixels on both sides; the geometry networks receive their gradient through the
    # antialias position gradient and the normals, under SMOOTH upstream gradients.
    assert worst["tex"] < 5e-3 and worst["dino"] < 5e-3 and worst["light"] < 5e-3, worst
    assert worst["sdf"] < 2e-2 and worst["arti"] < 2e-2, worst
    print("reference-caller chain: worst relative L2 gradient error per network", worst)


def test_articulation_constraints_installed_on_the_reference_class(ref, cuda):
    """predictors.install puts the one-kernel constraints on the reference's InstancePredictorBase; its own method, run on the same
    device tensors, is the checker (values and gradient), for the horse config and the bird config (`static_root_bones`)."""
    P = pkg("predictors")
    cls = ref.IPB.InstancePredictorBase
    original = cls.apply_articulation_constraints
    try:
        for kw in (dict(), dict(n_legs=0, n_leg_bones=0, mode="z_minmax")):
            pred = _instance_predictor(ref, cuda, **kw)
            if kw:
                pred.cfg_articulation.static_root_bones = True
            K = pred.num_bones
            torch.manual_seed(K)
            x, g = torch.randn(BATCH, 1, K, 3, device=cuda) * 8, torch.randn(BATCH, 1, K, 3, device=cuda)
            res = []
            for method in (original, P.apply_articulation_constraints):
                cls.apply_articulation_constraints = method
                xr = x.clone().requires_grad_(True)
                y = pred.apply_articulation_constraints(xr * 1.0)          # the reference method scales its argument in place
                y.backward(g)
                res.append((y.detach(), xr.grad))
            # torch's CUDA division by a scalar multiplies by the reciprocal: one more ulp than the CPU golden comparison
            assert float((res[0][0] - res[1][0]).abs().max()) <= 4e-7
            assert float((res[0][1] - res[1][1]).abs().max()) <= 2e-6 * float(res[0][1].abs().max())
    finally:
        cls.apply_articulation_constraints = original


def test_bird_chain_no_legs_static_root(ref, cuda):
    """train_magicpony_bird.yaml:29-36 - 8 body bones, no legs, `static_root_bones`: forward_articulation + its constraint masks
    through the drop-in; the posed mesh against the oracle."""
    base = _base_predictor(ref, cuda)
    pred = _instance_predictor(ref, cuda, n_body=8, n_legs=0, n_leg_bones=0, mode="z_minmax")
    pred.cfg_articulation.static_root_bones = True
    feat = torch.randn(BATCH, 32, device=cuda)
    patch = torch.randn(BATCH, 32, 8, 8, device=cuda)
    mvp, w2c, campos = _cameras(cuda, BATCH, seed=5)
    prior_shape, _ = base.forward(total_iter=0, is_training=False)
    shape, arti, aux = pred.forward_articulation(prior_shape, feat, patch, mvp, w2c, BATCH, 1, epoch=0, total_iter=10)
    assert float(arti[:, :, [3, 7]].abs().max()) == 0 and float(arti.abs().max()) > 0            # root bones are static
    verts = prior_shape.v_pos[0].detach().cpu()
    bones, chain, _ = gnp.estimate_bones(verts.numpy()[None, None], 8, n_legs=0, n_leg_bones=0, body_bones_mode="z_minmax")
    assert [(b, list(d)) for b, d in pred.kinematic_tree] == [(b, list(d)) for b, d in chain]
    posed, saux = T.skinning(verts[None, None], torch.from_numpy(bones), chain, arti.detach().cpu(), temperature=0.05)
    assert rel_err(shape.v_pos.detach().cpu().numpy(), posed[:, 0].numpy()) < 1e-4
    assert torch.equal(shape.t_pos_idx, prior_shape.t_pos_idx) and shape.v_pos.shape[0] == BATCH


def test_fauna_random_view_mask(ref, cuda):
    """FaunaModel.get_random_view_mask (Fauna.py:93-171): the second, texture-less / light-less ['shaded'] render of every
    3D-Fauna step, from cameras it derives itself - the reference's method on the drop-in, mask against the oracle."""
    if ref.Fauna is None:
        pytest.skip("Fauna.py did not import here: %s" % getattr(ref, "Fauna_error", ""))
    base = _base_predictor(ref, cuda)
    prior_shape, _ = base.forward(total_iter=0, is_training=False)
    pred = _instance_predictor(ref, cuda)
    feat = torch.randn(BATCH, 32, device=cuda)
    patch = torch.randn(BATCH, 32, 8, 8, device=cuda)
    mvp, w2c, campos = _cameras(cuda, BATCH)
    shape, arti, aux = pred.forward_articulation(prior_shape, feat, patch, mvp, w2c, BATCH, 1, epoch=0, total_iter=10)
    F_ = ref.Fauna
    model = RC.bare(F_.FaunaModel)
    model.cfg_render = ref.AM.RenderConfig(spatial_scale=SCALE, background_mode="none", renderer_spp=1, cam_pos_z_offset=10.0, fov=25.0)
    model.glctx = None
    model.accelerator = types.SimpleNamespace(device=cuda)
    torch.manual_seed(11)
    aux_out = model.get_random_view_mask(w2c, shape, prior_shape, 1, bins=360)
    mask = aux_out["mask_random_pred"]
    assert tuple(mask.shape) == (BATCH, 1, 256, 256)
    assert 0.005 < float(mask.mean()) < 0.6 and float(mask.min()) >= 0 and float(mask.max()) <= 1
    # the same cameras rebuilt from rand_degree (Fauna.py:113-142); mask = antialiased alpha of a texture-less, light-less render
    syn = pkg("synthetic")
    ang = (2 * np.pi / 360) * aux_out["rand_degree"].cpu().numpy().astype(np.float64)
    rot = np.stack([np.array([[np.cos(a), 0, np.sin(a), 0], [0, 1, 0, 0], [-np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]], np.float32) for a in ang])
    w2c_r = np.tile(np.eye(4, dtype=np.float32), (BATCH, 1, 1))
    w2c_r[:, :3, 3] = w2c.cpu().numpy()[:, :3, 3]
    proj = syn.perspective(25.0 / 180 * np.pi, 1.0, 0.1, 1000.0)
    mvp_r = torch.from_numpy(np.matmul(np.matmul(proj[None], w2c_r), rot).astype(np.float32))
    campos_r = torch.from_numpy(np.matmul(rot[:, :3, :3].transpose(0, 2, 1), -w2c_r[:, :3, 3][:, :, None])[:, :, 0].astype(np.float32))
    verts = shape.v_pos.detach().cpu()
    faces = shape.t_pos_idx[0].cpu()

    def shade(gb_tex, cam_normal, gbuf):
        kd = torch.ones(*gb_tex.shape[:-1], 3)
        return {"shaded": kd, "kd": kd}

    out = T.render_mesh(verts, T.auto_normals(verts, faces), faces, mvp_r, torch.from_numpy(w2c_r), campos_r, shade, (256, 256), spp=1,
                        background=torch.zeros(BATCH, 256, 256, 3), render_modes=("shaded",), prior_v_pos=prior_shape.v_pos.detach().cpu(),
                        two_sided_shading=False)
    want = out["shaded"][:, 3:].clamp(0, 1).numpy()
    got = mask.detach().cpu().numpy()
    bad = np.abs(got - want) > 1e-4
    assert bad.mean() < 2e-3, float(bad.mean())


def test_mesh_export_through_reference_save_obj(ref, cuda, tmp_path):
    """misc.save_obj -> write_obj -> material.save_mtl -> render_uv -> misc.save_images (model/utils/misc.py:170-187, obj.py:128-177,
    material.py:106-140): the reference's own export chain of the test / visualize configs, on the drop-in (write_obj text from
    libb2a.so, UV-atlas rasterisation + texture field on the device).  Files: .obj, .mtl, three texture PNGs."""
    import importlib
    import cv2
    material = importlib.import_module("model.render.material")
    base = _base_predictor(ref, cuda)
    prior_shape, _ = base.forward(total_iter=0, is_training=False)
    torch.manual_seed(4)
    mm = torch.tensor([[0., 1.]] * 9, device=cuda)
    tex = ref.networks.CoordMLP(3, 9, 4, nf=64, activation="sigmoid", min_max=mm, n_harmonic_functions=10, embedder_scalar=2 * np.pi / SCALE * 0.9,
                                extra_feat_dim=32, symmetrize=True).to(cuda)
    feat = torch.randn(1, 32, device=cuda)
    with torch.no_grad():
        mesh = prior_shape.clone()
        mesh.material = material.Material({"bsdf": "diffuse", "kd_ks_normal": tex})
        ref.misc.save_obj(str(tmp_path), meshes=mesh, save_material=True, feat=feat, fnames=["horse_mesh"], resolution=[128, 128])
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == ["horse_mesh.mtl", "horse_mesh.obj", "horse_texture_kd.png", "horse_texture_ks.png", "horse_texture_n.png"], names
    mtl = (tmp_path / "horse_mesh.mtl").read_text()
    assert mtl == "newmtl defaultMat\nbsdf   diffuse\nmap_Kd horse_texture_kd.png\nmap_Ks horse_texture_ks.png\nbump horse_texture_n.png\n"
    obj = (tmp_path / "horse_mesh.obj").read_text().splitlines()
    V, F = prior_shape.v_pos.shape[1], prior_shape.t_pos_idx.shape[1]
    assert obj[0] == "mtllib horse_mesh.mtl" and sum(l.startswith("v ") for l in obj) == V and sum(l.startswith("f ") for l in obj) == F
    # the kd texture = the texture field sampled at the UV atlas' world positions (render_uv), 8-bit, BGR on disk
    render = importlib.import_module("model.render.render")
    with torch.no_grad():
        mask, kd, ks, nrm = render.render_uv(None, mesh.get_n(0), [128, 128], tex, feat=feat)
    img = cv2.imread(str(tmp_path / "horse_texture_kd.png"), cv2.IMREAD_UNCHANGED)
    want = np.uint8(kd[0].cpu().numpy() * 255.0)[..., ::-1]
    assert img.shape == (128, 128, 3) and np.array_equal(img, want) and float(mask.mean()) > 0.001
