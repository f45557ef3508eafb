"""Generate tests/golden/*.npz by running the REFERENCE's own Python (build container only: needs /root/reference).

    python tests/golden/make_goldens.py

The reference ships no golden vectors or tests for the hot path, so these pin it: the reference's
`DMTet.__call__`, `estimate_bones`, `skinning` (+ autograd grads), `bsdf_prepare_shading_normal` and `DirectionalLight`, loaded by file
path on CPU tensors (oracle/reference_loader.py), on small seeded inputs.  The fixtures travel to the GPU box; the
reference tree does not.
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_loader  # noqa: E402

syn = importlib.import_module("3danimals_b200.synthetic")
OUT = os.path.dirname(os.path.abspath(__file__))


def mt_cases():
    ref = reference_loader.load()
    mt = reference_loader.reference_dmtet("cpu")
    for res in (12, 20):
        v, t = syn.kuhn_tet_grid(res)
        v = v * np.float32(7.0)
        for name, sdf in (("ellipsoid", syn.sdf_ellipsoid(v)), ("noisy_sphere", syn.sdf_noisy_sphere(v, 1.75, 0.05, 0)),
                          ("two_blobs", syn.sdf_two_blobs(v)), ("horse", syn.sdf_horse(v, 0.01, 0))):
            pos = torch.from_numpy(v)
            s = torch.from_numpy(sdf)[:, None].clone().requires_grad_(True)
            verts, faces, uvs, uv_idx = mt(pos, s, torch.from_numpy(t))
            g = torch.from_numpy(np.random.RandomState(5).randn(*verts.shape).astype(np.float32))
            (verts * g).sum().backward()
            np.savez_compressed(os.path.join(OUT, "mt_%s_%d.npz" % (name, res)), res=res, sdf=sdf, verts=verts.detach().numpy(),
                                faces=faces.numpy().astype(np.int32), uv_idx=uv_idx.numpy().astype(np.int64),
                                uvs_shape=np.array(uvs.shape), uvs_head=uvs[:64].numpy(), uvs_sum=uvs.double().sum().item(),
                                d_verts=g.numpy(), d_sdf=s.grad.numpy().reshape(-1))
    del ref


def skin_cases():
    ref = reference_loader.load()
    mt = reference_loader.reference_dmtet("cpu")
    v, t = syn.kuhn_tet_grid(20)
    v = v * np.float32(7.0)
    verts, faces, _, _ = mt(torch.from_numpy(v), torch.from_numpy(syn.sdf_horse(v, 0.0, 0))[:, None], torch.from_numpy(t))
    verts = verts.detach()
    for name, n_leg, mode, B in (("horse", 3, "z_minmax_y+", 2), ("bird", 0, "z_minmax", 3)):
        bones, chain, aux = ref.skinning.estimate_bones(verts[None, None], 8, n_legs=4, n_leg_bones=n_leg, body_bones_mode=mode,
                                                        compute_kinematic_chain=True)
        bones2 = ref.skinning.estimate_bones(verts[None, None] * 1.01, 8, n_legs=4, n_leg_bones=n_leg, body_bones_mode=mode,
                                             compute_kinematic_chain=False, aux=aux)
        K = bones.shape[2]
        rng = np.random.RandomState(11)
        ang = torch.from_numpy(rng.uniform(-0.5, 0.5, (B, 1, K, 3)).astype(np.float32)).requires_grad_(True)
        vp = verts[None, None].clone().requires_grad_(True)
        out, saux = ref.skinning.skinning(vp, bones, chain, ang, output_posed_bones=True, temperature=0.05)
        g = torch.from_numpy(rng.randn(*out.shape).astype(np.float32))
        gp = torch.from_numpy(rng.randn(*saux["posed_bones"].shape).astype(np.float32))
        ((out * g).sum() + (saux["posed_bones"] * gp).sum()).backward()
        np.savez_compressed(os.path.join(OUT, "skin_%s.npz" % name), verts=verts.numpy(), faces=faces.numpy().astype(np.int32),
                            n_leg_bones=n_leg, mode=mode, bones=bones.numpy(), bones_rescaled=bones2.numpy(),
                            chain_ids=np.array([b for b, _ in chain]), chain_dep=np.array([",".join(map(str, d)) for _, d in chain]),
                            angles=ang.detach().numpy(), out=out.detach().numpy(), posed_bones=saux["posed_bones"].detach().numpy(),
                            weights=saux["vertices_to_bones"].detach().numpy(), g_out=g.numpy(), g_posed=gp.numpy(),
                            d_angles=ang.grad.numpy(), d_verts=vp.grad.numpy())


def shading_case():
    ref = reference_loader.load()
    rng = np.random.RandomState(21)
    shp = (2, 6, 5, 3)
    t = lambda: torch.from_numpy(rng.randn(*shp).astype(np.float32))
    pos, nrm, tng, geo = t(), t(), t(), t()
    view = torch.from_numpy(rng.randn(2, 1, 1, 3).astype(np.float32) * 3)
    for x in (pos, nrm, geo):
        x.requires_grad_(True)
    pert = torch.tensor([0, 0, 1], dtype=torch.float32)[None, None, None]
    out = ref.bsdf.bsdf_prepare_shading_normal(pos, view, pert, nrm, tng, geo, True, True)
    g = t()
    (out * g).sum().backward()
    np.savez_compressed(os.path.join(OUT, "shading_normal.npz"), pos=pos.detach().numpy(), view=view.numpy(), nrm=nrm.detach().numpy(),
                        tng=tng.numpy(), geo=geo.detach().numpy(), out=out.detach().numpy(), g=g.numpy(), d_pos=pos.grad.numpy(),
                        d_nrm=nrm.grad.numpy(), d_geo=geo.grad.numpy())


def fauna_bones_case():
    """3D-Fauna's estimate_bones variant (InstancePredictorFauna.py:20,84-99: bone_y_threshold = 0.4, seven masked
    quantiles, skinning.py:163-175), single shape and a batch of per-instance shapes."""
    ref = reference_loader.load()
    mt = reference_loader.reference_dmtet("cpu")
    v, t = syn.kuhn_tet_grid(20)
    v = v * np.float32(7.0)
    verts, _, _, _ = mt(torch.from_numpy(v), torch.from_numpy(syn.sdf_horse(v, 0.0, 0))[:, None], torch.from_numpy(t))
    verts = verts.detach()
    rng = np.random.RandomState(17)
    batch = torch.stack([verts * (1 + 0.04 * i) + torch.from_numpy(rng.randn(*verts.shape).astype(np.float32)) * 0.01 for i in range(3)])[:, None]
    out = {}
    for tag, shape in (("single", verts[None, None]), ("batch", batch)):
        bones, chain, aux = ref.skinning.estimate_bones(shape, 8, n_legs=4, n_leg_bones=3, body_bones_mode="z_minmax_y+",
                                                        compute_kinematic_chain=True, bone_y_threshold=0.4)
        bones2 = ref.skinning.estimate_bones(shape * 1.01, 8, n_legs=4, n_leg_bones=3, body_bones_mode="z_minmax_y+",
                                             compute_kinematic_chain=False, aux=aux, bone_y_threshold=0.4)
        out.update({tag + "_shape": shape.numpy(), tag + "_bones": bones.numpy(), tag + "_bones_rescaled": bones2.numpy(),
                    tag + "_chain_ids": np.array([b for b, _ in chain]), tag + "_chain_dep": np.array([",".join(map(str, d)) for _, d in chain]),
                    tag + "_attach": np.array([l["body_bone_idx"] for l in aux["legs"]])})
    np.savez_compressed(os.path.join(OUT, "bones_fauna.npz"), **out)


def light_case():
    """The reference's DirectionalLight (light.py:168-193): light MLP -> light_params, shade(feat, kd, normal) + grads."""
    ref = reference_loader.load()
    torch.manual_seed(5)
    lgt = ref.light.DirectionalLight(16, 3, 32, intensity_min_max=torch.FloatTensor([[0.0, 1.0], [0.5, 1.0]]))
    rng = np.random.RandomState(22)
    B, H, W = 3, 6, 10
    feat = torch.from_numpy(rng.randn(B, 16).astype(np.float32)).requires_grad_(True)
    tex = torch.from_numpy(rng.rand(B, H, W, 9).astype(np.float32)).requires_grad_(True)     # kd = leading 3 of 9 channels
    nrm = rng.randn(B, H, W, 3).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    nrm[:, 0] = 0                                                                            # uncovered pixels: zero normal (dot == 0)
    nrm = torch.from_numpy(nrm).requires_grad_(True)
    shaded, shading = lgt.shade(feat, tex[..., :3], nrm)
    g1 = torch.from_numpy(rng.randn(*shaded.shape).astype(np.float32))
    g2 = torch.from_numpy(rng.randn(*shading.shape).astype(np.float32))
    ((shaded * g1).sum() + (shading * g2).sum()).backward()
    lp = lgt.light_params.detach()
    sd = {k: v.detach().numpy() for k, v in lgt.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "light_directional.npz"), feat=feat.detach().numpy(), tex=tex.detach().numpy(), nrm=nrm.detach().numpy(),
                        light_params=lp.numpy(), shaded=shaded.detach().numpy(), shading=shading.detach().numpy(), g_shaded=g1.numpy(),
                        g_shading=g2.numpy(), d_feat=feat.grad.numpy(), d_tex=tex.grad.numpy(), d_nrm=nrm.grad.numpy(),
                        **{"sd:" + k: v for k, v in sd.items()})


def obj_case():
    """The reference's own write_obj (obj.py:128-177) on a small mesh whose coordinates cover the special values of the number
    format, in the three layouts the callers produce: full (texcoords + normals), save_material=False (no 'vt' lines, texcoord
    column kept) and bare (no texcoords, no normals).  material=None so no .mtl is written (obj.py:169)."""
    import contextlib
    import io
    import tempfile
    import types
    from oracle import obj_text as oracle_obj
    write_obj = reference_loader.reference_write_obj()
    rng = np.random.RandomState(31)
    sp = oracle_obj.special_float32()
    n = (len(sp) + 2) // 3
    v_pos = np.concatenate([sp, rng.randn(3 * n - len(sp)).astype(np.float32)]).reshape(n, 3)
    bits = rng.randint(0, 2 ** 32, size=(200, 3), dtype=np.uint64).astype(np.uint32).view(np.float32)      # any bit pattern
    v_pos = np.concatenate([v_pos, bits, (rng.randn(100, 3) * 3).astype(np.float32)])
    V = len(v_pos)
    v_nrm = rng.randn(V, 3).astype(np.float32)
    v_nrm /= np.linalg.norm(v_nrm, axis=-1, keepdims=True)
    v_tex = np.concatenate([rng.rand(150, 2), sp[:60].reshape(30, 2)]).astype(np.float32)
    F = 400
    t_pos = rng.randint(0, V, size=(F, 3)).astype(np.int64)
    t_tex = rng.randint(0, len(v_tex), size=(F, 3)).astype(np.int64)
    out = dict(v_pos=v_pos, v_nrm=v_nrm, v_tex=v_tex, t_pos=t_pos, t_tex=t_tex)
    T = lambda a: None if a is None else torch.from_numpy(a)[None]
    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        for key, (nrm, tex, save_material) in dict(full=(v_nrm, v_tex, True), nomat=(v_nrm, v_tex, False), bare=(None, None, True)).items():
            m = types.SimpleNamespace(v_pos=T(v_pos), v_nrm=T(nrm), v_tex=T(tex), t_pos_idx=T(t_pos), t_nrm_idx=T(t_pos) if nrm is not None else None,
                                      t_tex_idx=T(t_tex) if tex is not None else None, material=None)
            write_obj(d, "golden_" + key, m, 0, save_material=save_material)
            out["text_" + key] = np.frombuffer(open(os.path.join(d, "golden_" + key + ".obj"), "rb").read(), np.uint8)
    np.savez_compressed(os.path.join(OUT, "obj_export.npz"), **out)


def geometry_module_case():
    """The torch side of R1 (DMTetGeometry, dmtet.py:161-281): the reference class itself on CPU - its CUDA-only grid load is
    replaced by an in-memory Kuhn grid - with a seeded SDF MLP: get_sdf for every init_sdf mode (+ symmetrize), the BCE edge
    regulariser on the unique sorted edge list, the eikonal sample gradients and both regulariser values under a fixed RNG seed."""
    ref = reference_loader.load()
    D = ref.dmtet
    v, t = syn.kuhn_tet_grid(6)
    out = dict(verts=v, tets=t.astype(np.int64))
    orig_load, orig_dmtet = D.DMTetGeometry.load_tets, D.DMTet

    def load_tets(self, grid_res=None, scale=None):
        self.verts = torch.from_numpy(v) * self.grid_scale
        self.indices = torch.from_numpy(t.astype(np.int64))
        e = self.indices[:, torch.tensor([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3])].reshape(-1, 2)      # generate_edges (:283-288) on CPU
        self.all_edges = torch.unique(torch.sort(e, dim=1)[0], dim=0)

    D.DMTetGeometry.load_tets, D.DMTet = load_tets, (lambda device=None: orig_dmtet(device="cpu"))
    try:
        torch.manual_seed(41)
        geo = D.DMTetGeometry(6, 7.0, num_layers=5, hidden_size=32, embedder_freq=8, embed_concat_pts=True, init_sdf="ellipsoid",
                              jitter_grid=0.05, symmetrize=True)
    finally:
        D.DMTetGeometry.load_tets, D.DMTet = orig_load, orig_dmtet
    with torch.no_grad():
        for p_ in geo.mlp.parameters():
            p_.mul_(3.0)                                    # away from the near-zero default init
    for k, w in geo.mlp.state_dict().items():
        out["sd:" + k] = w.numpy()
    out["all_edges"] = geo.all_edges.numpy()
    pts = (torch.rand(300, 3) - 0.5) * 7.0
    out["pts"] = pts.numpy()
    for mode in ("ellipsoid", "sphere", 0.25, None):
        geo.init_sdf = mode
        out["sdf_%s" % mode] = geo.get_sdf(pts).detach().numpy()
        out["sdf_grid_%s" % mode] = geo.get_sdf().detach().numpy()
    geo.init_sdf = "ellipsoid"
    geo.symmetrize = False
    out["sdf_nosym"] = geo.get_sdf(pts).detach().numpy()
    geo.symmetrize = True
    geo.current_sdf = geo.get_sdf()
    out["bce"] = D.sdf_bce_reg_loss(geo.current_sdf, geo.all_edges).detach().numpy()
    geo.mesh_verts = (torch.rand(6000, 3) - 0.5) * 2.0
    out["mesh_verts"] = geo.mesh_verts.numpy()
    torch.manual_seed(43)
    out["eikonal_grad"] = geo.get_sdf_gradient().detach().numpy()
    torch.manual_seed(43)
    reg = geo.get_sdf_reg_loss()
    out["reg_bce"], out["reg_grad"] = reg["sdf_bce_reg_loss"].detach().numpy(), reg["sdf_gradient_reg_loss"].detach().numpy()
    (reg["sdf_bce_reg_loss"] + reg["sdf_gradient_reg_loss"]).backward()         # double backward through the eikonal term
    for k, p_ in geo.mlp.named_parameters():
        out["grad:" + k] = p_.grad.numpy()
    lo, hi = geo.getAABB()
    out["aabb"] = torch.stack([lo, hi]).numpy()
    np.savez_compressed(os.path.join(OUT, "dmtet_geometry.npz"), **out)


def normals_case():
    """R3: the reference's own `auto_normals` (model/render/mesh.py:276-304) on CPU tensors.  The function hard-codes
    device='cuda' for its fallback constant (:297-298); the file is executed unmodified and `torch.tensor` is intercepted for
    the duration of the call so that literal lands on the CPU.  Cases: an extracted DMTet mesh (batch 2, second instance
    perturbed), plus a mesh with a vertex no face uses and a vertex touched only by zero-area faces (fallback (0,0,1))."""
    ref = reference_loader.load()
    mt = reference_loader.reference_dmtet("cpu")
    v, t = syn.kuhn_tet_grid(12)
    v = v * np.float32(7.0)
    verts, faces, _, _ = mt(torch.from_numpy(v), torch.from_numpy(syn.sdf_horse(v, 0.01, 0))[:, None], torch.from_numpy(t))
    verts = verts.detach()
    rng = np.random.RandomState(21)
    real_tensor = torch.tensor

    def cpu_tensor(*a, **k):
        if k.get("device") == "cuda":
            k["device"] = "cpu"
        return real_tensor(*a, **k)

    out = {}
    deg_v = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 1], [2, 0, 0], [5, 5, 5], [3, 0, 0]], dtype=torch.float32)   # 4: unused; 5: only in a collinear face
    deg_f = torch.tensor([[0, 1, 2], [1, 3, 5]], dtype=torch.long)
    for name, vp, fc in (("mesh", torch.stack([verts, verts + torch.from_numpy(rng.randn(*verts.shape).astype(np.float32)) * 0.02]), faces),
                         ("degenerate", deg_v[None], deg_f)):
        vp = vp.clone().requires_grad_(True)
        m = ref.mesh.Mesh(vp, fc[None])
        torch.tensor = cpu_tensor
        try:
            nm = ref.mesh.auto_normals(m)
        finally:
            torch.tensor = real_tensor
        g = torch.from_numpy(rng.randn(*nm.v_nrm.shape).astype(np.float32))
        (nm.v_nrm * g).sum().backward()
        assert nm.t_nrm_idx is m.t_pos_idx or torch.equal(nm.t_nrm_idx, m.t_pos_idx)
        out.update({name + "_v_pos": vp.detach().numpy(), name + "_faces": fc.numpy().astype(np.int32), name + "_v_nrm": nm.v_nrm.detach().numpy(),
                    name + "_g": g.numpy(), name + "_d_v_pos": vp.grad.numpy()})
    np.savez_compressed(os.path.join(OUT, "normals.npz"), **out)


def raster_case():
    """REGRESSION PINS of the rasterizer restatement (oracle/raster_ref.c), not reference outputs: nvdiffrast is absent, so
    nothing upstream can be run (DESIGN.md §2 "parity unpinned").  They freeze today's ids / barycentrics / interpolation /
    antialiasing / gradients on the analytic scenes (tests/raster_scenes.py) and one extracted mesh, so that a later edit of the
    restatement (the yardstick of every GPU parity test) cannot drift silently."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from raster_scenes import ANALYTIC, RESOLUTIONS
    from oracle import geometry_np as gnp
    from oracle import raster as R
    out = {}
    scenes = {k: v for k, v in ANALYTIC.items()}
    v, t = syn.kuhn_tet_grid(16)
    v = v * np.float32(7.0)
    o = gnp.marching_tets(v, syn.sdf_horse(v, 0.01, 3), t, with_uvs=False)
    mvp, _, _ = syn.cameras(2, seed=9)
    scenes["horse_mesh"] = (R.xfm_points(o["verts"][None], mvp), o["faces"].astype(np.int32))
    for name, (pos, tri) in sorted(scenes.items()):
        for res in (RESOLUTIONS if name != "horse_mesh" else [(64, 64)]):
            key = "%s:%dx%d:" % (name, res[0], res[1])
            rng = np.random.RandomState(len(key) * 7 + res[0])
            rast = R.rasterize(pos, tri, res)
            attr = rng.randn(pos.shape[0], pos.shape[1], 5).astype(np.float32)
            col = R.interpolate(attr, rast, tri)
            aa = R.antialias(col, rast, pos, tri)
            g = rng.randn(*col.shape).astype(np.float32)
            d_attr, d_rast = R.interpolate_bwd(attr, rast, tri, g)
            d_col, d_pos_aa = R.antialias_bwd(col, rast, pos, tri, g)
            d_pos_r = R.rasterize_bwd(pos, tri, rast, d_rast)
            if name == "horse_mesh":
                out[key + "pos"], out[key + "tri"] = pos, tri
            for k, a in dict(rast=rast, attr=attr, col=col, aa=aa, g=g, d_attr=d_attr, d_rast=d_rast, d_col=d_col, d_pos_aa=d_pos_aa, d_pos_r=d_pos_r).items():
                out[key + k] = a
    np.savez_compressed(os.path.join(OUT, "raster_scenes.npz"), **out)


def _reference_methods(path, names):
    """The named methods of the first class in a reference source file, compiled on their own (the module imports a ViT stack)."""
    import ast
    tree = ast.parse(open(path).read())
    ns = {"torch": torch, "np": np, "print": lambda *a, **k: None}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), ns)
    return [ns[n] for n in names]


def articulation_configs():
    """(name, cfg_articulation, cfg_additional or None, total_iter): the shipped configs (config/model/*.yaml, config/*.yaml overrides)
    plus the Fauna-constraint and late-iteration branches no shipped config enables."""
    from types import SimpleNamespace as NS
    horse = dict(num_body_bones=8, num_leg_bones=3, num_legs=4, output_multiplier=0.1, static_root_bones=False, constrain_legs=True,
                 use_fauna_constraints=False, extra_constraints=False, max_arti_angle=60)
    add = dict(iter_leg_rotation_start=300000, forbid_leg_rotate=True, small_leg_angle=True, reg_body_rotate_mult=0.1)
    return [("magicpony_horse", NS(**horse), None, 0),
            ("magicpony_bird", NS(**dict(horse, num_leg_bones=0, static_root_bones=True, max_arti_angle=45)), None, 0),
            ("ponymation", NS(**dict(horse, extra_constraints=True)), None, 0),
            ("base_fauna_constraints", NS(**dict(horse, use_fauna_constraints=True, static_root_bones=True)), None, 0),
            ("fauna_early", NS(**dict(horse, constrain_legs=False)), NS(**add), 1000),
            ("fauna_late", NS(**dict(horse, constrain_legs=False, static_root_bones=True)), NS(**add), 300001),
            ("fauna_late_large_legs", NS(**dict(horse, constrain_legs=False)), NS(**dict(add, small_leg_angle=False)), 400000)]


def articulation_case():
    """InstancePredictorBase.apply_articulation_constraints and Fauna's sequence, the reference's own methods on CPU tensors."""
    from types import SimpleNamespace as NS
    root = os.path.join(reference_loader.REFERENCE_ROOT, "model", "predictors")
    (base,) = _reference_methods(os.path.join(root, "InstancePredictorBase.py"), ["apply_articulation_constraints"])
    f_con, f_reg = _reference_methods(os.path.join(root, "InstancePredictorFauna.py"), ["apply_articulation_constraints", "apply_fauna_articulation_regularizer"])
    out = {}
    for name, cfg, add, it in articulation_configs():
        K = cfg.num_body_bones + cfg.num_leg_bones * cfg.num_legs
        rng = np.random.RandomState(len(name))
        x = (rng.randn(3, 2, K, 3) * 8).astype(np.float32)
        g = rng.randn(3, 2, K, 3).astype(np.float32)
        xt = torch.from_numpy(x.copy()).requires_grad_(True)
        self = NS(cfg_articulation=cfg, cfg_additional=add)
        if add is None:
            y = base(self, xt * 1.0)                    # the method scales its argument in place: hand it a non-leaf copy
        else:
            a = xt * 1.0
            a *= cfg.output_multiplier                  # InstancePredictorFauna.py:228-229
            a = a.tanh()
            y = f_reg(self, f_con(self, a, it), it)
        y.backward(torch.from_numpy(g))
        out[name + ":x"], out[name + ":g"], out[name + ":y"], out[name + ":d_x"] = x, g, y.detach().numpy(), xt.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "articulation.npz"), **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    if "--only-new" not in sys.argv:
        mt_cases()
        skin_cases()
        shading_case()
        light_case()
        fauna_bones_case()
    obj_case()
    articulation_case()
    normals_case()
    raster_case()
    geometry_module_case()
    print(sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))
