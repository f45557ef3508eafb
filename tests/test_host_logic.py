"""Host-side logic that needs no GPU: the torch restatement of estimate_bones (runs on CPU tensors), kinematic-chain
tables, the module overlay, synthetic inputs, Mesh bookkeeping."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import golden, golden_files, pkg
from oracle import geometry_np as gnp


def _chain(g):
    return [(int(b), [int(x) for x in str(d).split(",") if x != ""]) for b, d in zip(g["chain_ids"], g["chain_dep"])]


@pytest.mark.parametrize("name", golden_files("skin_"))
def test_estimate_bones_matches_reference_golden(name):
    """The torch formulation behind estimate_bones (the path Fauna's variant takes on the device) is device-agnostic: pin it
    on CPU tensors against the reference's output.  The public entry point itself refuses CPU tensors (no CPU fallback)."""
    sk = pkg("geometry.skinning")
    g = golden(name)
    verts = torch.from_numpy(g["verts"])[None, None]
    n_leg, mode = int(g["n_leg_bones"]), str(g["mode"])
    with pytest.raises(RuntimeError):
        sk.estimate_bones(verts, 8, n_legs=4, n_leg_bones=n_leg, body_bones_mode=mode)
    with torch.no_grad():
        bones, chain, aux = sk._estimate_bones_torch(verts, 8, n_leg, mode, True, None, True, None, None)
        assert [(b, list(d)) for b, d in chain] == _chain(g)
        assert np.allclose(bones.numpy(), g["bones"], atol=1e-6)
        bones2 = sk._estimate_bones_torch(verts * 1.01, 8, n_leg, mode, False, aux, True, None, None)
        assert np.allclose(bones2.numpy(), g["bones_rescaled"], atol=1e-6)


def test_estimate_bones_batched_matches_numpy_oracle():
    """B x F > 1 (per-instance deformation on, InstancePredictorBase.py:514-518): vectorised masked arg-min vs the oracle's
    per-(b,f) loops, including the 'attachment index fixed on the first (b,f)' behaviour (skinning.py:190-192)."""
    sk = pkg("geometry.skinning")
    g = golden("skin_horse.npz")
    rng = np.random.RandomState(0)
    shapes = np.stack([g["verts"] * np.float32(1 + 0.05 * i) + rng.randn(*g["verts"].shape).astype(np.float32) * 0.01
                       for i in range(6)]).reshape(3, 2, -1, 3)
    ref_b, ref_chain, ref_aux = gnp.estimate_bones(shapes, 8, n_legs=4, n_leg_bones=3, body_bones_mode="z_minmax_y+")
    with torch.no_grad():
        bones, chain, aux = sk._estimate_bones_torch(torch.from_numpy(shapes), 8, 3, "z_minmax_y+", True, None, True, None, None)
    assert [(b, list(d)) for b, d in chain] == [(b, list(d)) for b, d in ref_chain]
    assert np.allclose(bones.numpy(), ref_b, atol=1e-6)
    assert [l["body_bone_idx"] for l in aux["legs"]] == [l["body_bone_idx"] for l in ref_aux["legs"]]


def test_euler_and_chain_tables():
    sk, ops = pkg("geometry.skinning"), pkg("ops")
    a = torch.tensor([[0.3, -0.2, 0.5]])
    assert np.allclose(sk.euler_angles_to_matrix(a, "XYZ").numpy(), gnp.euler_xyz(a.numpy()), atol=1e-6)
    with pytest.raises(ValueError):
        sk.euler_angles_to_matrix(a, "XXY")
    chain = [(2, [0, 1]), (1, [0]), (0, []), (3, [])]
    ptr, ids = ops.chain_tables(chain, 4, "cpu")
    # bone 0's ancestors are the bones listing it among their dependents, in list order, root first (skinning.py:389-396)
    assert ptr.tolist() == [0, 3, 5, 6, 7] and ids.tolist() == [2, 1, 0, 2, 1, 2, 3]
    assert gnp.chain_lists(chain) == {2: [2], 1: [2, 1], 0: [2, 1, 0], 3: [3]}


def test_synthetic_grid_schema(tmp_path):
    syn = pkg("synthetic")
    v, t = syn.kuhn_tet_grid(4)
    assert v.shape == (125, 3) and t.shape == (6 * 64, 4) and v.dtype == np.float32 and t.dtype == np.int64
    # every tet has positive volume magnitude and the 6 tets tile each cube exactly
    p = v[t]
    vol = np.abs(np.einsum("ni,ni->n", np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), p[:, 3] - p[:, 0])) / 6
    assert np.allclose(vol.sum(), 1.0, atol=1e-5) and vol.min() > 0
    path = syn.write_tet_npz(4, str(tmp_path))
    z = np.load(path)
    assert set(z.files) == {"vertices", "indices"}          # the reference's npz schema (dmtet.py:223-225)
    mvp, w2c, campos = syn.cameras(3)
    assert mvp.shape == (3, 4, 4) and np.allclose(mvp, syn.perspective(25 / 180 * np.pi, 1.0, 0.1, 1000.0) @ w2c, atol=1e-5)
    assert np.allclose(campos, -np.einsum("bji,bj->bi", w2c[:, :3, :3], w2c[:, :3, 3]), atol=1e-5)


def test_overlay_aliases_reference_module_names():
    ov = pkg("overlay")
    ov.install()
    try:
        import nvdiffrast.torch as dr
        assert dr.__name__ == "3danimals_b200.nvdiffrast_shim.torch"
        assert hasattr(dr, "RasterizeGLContext") and hasattr(dr, "DepthPeeler") and hasattr(dr, "antialias")
        for alias, target in ov.ALIASES.items():
            if alias.startswith("model."):
                # parent packages come from the reference tree; resolve the alias through the finder directly
                spec = ov._finder.find_spec(alias)
                assert spec is not None and spec.loader.create_module(spec).__name__ == target
        with pytest.raises(NotImplementedError):
            dr.texture()
    finally:
        ov.uninstall()
    assert "nvdiffrast.torch" not in sys.modules


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference tree only exists in the build container")
def test_overlay_lets_reference_predictor_import_our_geometry():
    """A reference file that imports the hot path (`from ..geometry.dmtet import DMTetGeometry`, BasePredictorBase.py:10)
    binds to the B200 modules once the overlay is installed, with the rest of `model.*` still from the reference."""
    import types
    ov = pkg("overlay")
    saved = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.")}
    ov.install()
    try:
        for name, rel in (("model", "model"), ("model.geometry", "model/geometry"), ("model.render", "model/render"),
                          ("model.networks", "model/networks"), ("model.predictors", "model/predictors"), ("model.utils", "model/utils")):
            m = types.ModuleType(name)
            m.__path__ = [os.path.join("/root/reference", rel)]
            sys.modules[name] = m
        for stub in ("imageio", "omegaconf", "omegaconf.errors"):
            sys.modules.setdefault(stub, types.ModuleType(stub))
        sys.modules["omegaconf.errors"].ConfigAttributeError = AttributeError
        mod = importlib.import_module("model.predictors.BasePredictorBase")
        assert mod.DMTetGeometry.__module__ == "3danimals_b200.geometry.dmtet"
        assert importlib.import_module("model.render.render").__name__ == "3danimals_b200.render.render"
        assert importlib.import_module("model.networks.MLPs").__file__.startswith("/root/reference")
        # model.render.light: DirectionalLight is the fused-shade drop-in (built on the reference's own MLP class), every
        # other name of the reference module is re-exported from the reference file
        mlps = importlib.import_module("model.networks.MLPs")
        for sym in ("MLP", "CoordMLP"):      # what the reference's model/networks/__init__.py exports (the skeleton above skips it)
            setattr(sys.modules["model.networks"], sym, getattr(mlps, sym))
        light = importlib.import_module("model.render.light")
        assert light.DirectionalLight.__module__ == "3danimals_b200.render.light"
        assert hasattr(light, "EnvironmentLight") and light.EnvironmentLight.__module__ == "model.render._reference_light"
        # model.render.obj: write_obj is the libb2a.so writer (what model/utils/misc.py:12 imports)
        assert importlib.import_module("model.render.obj").write_obj.__module__ == "3danimals_b200.render.obj"
        lgt = light.DirectionalLight(16, 3, 32, intensity_min_max=torch.zeros(2, 2))
        assert type(lgt.mlp).__module__ == "model.networks.MLPs" and sorted(lgt.state_dict()) == [
            "intensity_min_max", "mlp.network.0.weight", "mlp.network.2.weight", "mlp.network.4.weight"]
    finally:
        ov.uninstall()
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_mesh_bookkeeping_cpu():
    """Mesh container logic that involves no kernel: lazy UV broadcast, copy_none, clone."""
    mesh = pkg("render.mesh")
    v = torch.rand(3, 5, 3)
    f = torch.tensor([[[0, 1, 2], [2, 3, 4]]])
    uv = torch.rand(1, 8, 2)
    m = mesh.Mesh(v, f, v_tex=uv, t_tex_idx=f)
    assert m.v_tex.shape == (3, 8, 2) and m.v_tex.data_ptr() == uv.data_ptr()      # expand, not repeat
    c = m.clone()
    assert c.v_pos is not m.v_pos and torch.equal(c.v_pos, m.v_pos) and len(c) == 3
    assert mesh.Mesh(base=m).t_pos_idx is f
    assert torch.equal(mesh.compute_edges(f), torch.tensor([[0, 1], [0, 2], [1, 2], [2, 3], [2, 4], [3, 4]]))
    assert m.tri_i32().dtype == torch.int32


def test_directional_light_module_matches_reference_golden():
    """The drop-in DirectionalLight keeps the reference's parameter names (checkpoints load strictly) and its light MLP
    -> light_params arithmetic (light.py:177-184); the per-pixel shading has no CPU path."""
    import torch
    light_mod = pkg("render.light")
    g = golden("light_directional.npz")
    lgt = light_mod.DirectionalLight(16, 3, 32, intensity_min_max=torch.zeros(2, 2))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
    lgt.load_state_dict(sd, strict=True)
    lp = lgt(torch.from_numpy(g["feat"]))
    assert np.abs(lp.detach().numpy() - g["light_params"]).max() < 1e-6
    with pytest.raises(RuntimeError):
        lgt.shade(torch.from_numpy(g["feat"]), torch.from_numpy(g["tex"])[..., :3], torch.from_numpy(g["nrm"]))


@pytest.mark.parametrize("tag", ["single", "batch"])
def test_fauna_bones_variant_host_logic(tag):
    """The drop-in estimate_bones' bone_y_threshold branch (device-side torch formulation; runs on any device) vs the
    reference's golden: bones, kinematic chain, leg attachment, and the no-chain re-estimation of the next iteration."""
    import torch
    sk = pkg("geometry.skinning")
    g = golden("bones_fauna.npz")
    shape = torch.from_numpy(g[tag + "_shape"])
    with torch.no_grad():
        bones, chain, aux = sk._estimate_bones_torch(shape, 8, 3, "z_minmax_y+", True, None, True, None, 0.4)
        ref_chain = [(int(b), [int(x) for x in str(d).split(",") if x != ""]) for b, d in zip(g[tag + "_chain_ids"], g[tag + "_chain_dep"])]
        assert [(int(b), [int(x) for x in d]) for b, d in chain] == ref_chain
        assert [int(l["body_bone_idx"]) for l in aux["legs"]] == list(g[tag + "_attach"])
        assert np.allclose(bones.numpy(), g[tag + "_bones"], atol=1e-5)
        bones2 = sk._estimate_bones_torch(shape * 1.01, 8, 3, "z_minmax_y+", False, aux, True, None, 0.4)
        assert np.allclose(bones2.numpy(), g[tag + "_bones_rescaled"], atol=1e-5)


def test_capture_helpers_refuse_cpu_tensors():
    """CUDA-graph capture (3danimals_b200/graphs.py) is a device-only facility: CPU tensors are rejected up front."""
    import torch
    graphs = pkg("graphs")
    with pytest.raises(RuntimeError):
        graphs.CapturedStep(lambda x: x * 2, [torch.zeros(3)])
    ops = pkg("ops")
    assert not ops.composite_up_supported(torch.zeros(1, 4, 4, 3), None)          # no prepared context -> reference-order torch path
    assert not ops.pair_supported(torch.zeros(1, 4, 4, 16), torch.zeros(1, 4, 4, 3), None)


@pytest.mark.parametrize("which", ["twin", "reference"])
def test_field_feature_enters_as_per_image_bias(which):
    """Sparse field evaluation (§8f-1): the per-image feature half of CoordMLP's first hidden layer (MLPs.py:90-94) is applied as
    a per-image bias (`render._coord_mlp_rows`).  Against the module's own every-pixel forward - this package's twin and, when
    the tree is present, the reference's own class - outputs on covered pixels and every gradient (positions, feature,
    parameters) agree to fp32 summation order; pure PyTorch, so the CPU run covers the arithmetic of the device path."""
    from oracle import reference_loader
    R = pkg("render.render")
    if which == "reference":
        if not reference_loader.available():
            pytest.skip("reference tree only exists in the build container")
        cls = reference_loader.load().mlps.CoordMLP
    else:
        cls = pkg("networks").CoordMLP
    torch.manual_seed(3)
    cases = (dict(cin=3, cout=9, num_layers=8, nf=64, activation="sigmoid", min_max=torch.tensor([[0.0, 1.0]] * 6 + [[-1.0, 1.0]] * 3), extra_feat_dim=32, symmetrize=True),
             dict(cin=3, cout=16, num_layers=5, nf=48, activation="sigmoid", extra_feat_dim=24, in_layer_relu=True, n_harmonic_functions=0),
             dict(cin=3, cout=4, num_layers=2, nf=32, extra_feat_dim=8, embed_concat_pts=False))
    for kw in cases:
        net = cls(**kw)
        assert type(net).__name__ == "CoordMLP"
        B, h, w, C = 3, 9, 11, kw["extra_feat_dim"]
        gb = torch.randn(B, h, w, 3)
        feat = torch.randn(B, C, requires_grad=True)
        rast = torch.zeros(B, h, w, 4)
        rast[..., 3] = (torch.rand(B, h, w) > 0.4).float()
        rast[1, ..., 3] = 0                                     # an image without coverage
        sparse = R._covered_rows(rast)
        assert R._splits_feat(net, feat) and not R._splits_feat(net, None) and not R._splits_feat(torch.nn.Linear(3, 3), feat)
        up = torch.randn(B, h, w, kw["cout"])
        res = []
        for split in (True, False, None):                       # per-image bias / [N,C] feature rows / the reference's dense call
            net.zero_grad()
            feat.grad = None
            x = gb.clone().requires_grad_(True)
            saved = R.SPLIT_FIELD_FEAT
            try:
                R.SPLIT_FIELD_FEAT = bool(split)
                y = R._sample_field(net, x, feat, sparse if split is not None else None)
            finally:
                R.SPLIT_FIELD_FEAT = saved
            y = y * (rast[..., 3:] > 0)
            (y * up).sum().backward()
            res.append([y.detach(), x.grad.clone(), feat.grad.clone()] + [p.grad.clone() for p in net.parameters()])
        for other in res[1:]:
            for a, b in zip(res[0], other):
                assert float((a - b).abs().max()) <= 2e-5 * max(1.0, float(b.abs().max()))
    # a feature shared by the batch ([1,C]) and a module without features keep working
    net = cls(cin=3, cout=4, num_layers=3, nf=32, extra_feat_dim=8)
    y1 = R._sample_field(net, gb, feat[:1, :8].detach(), sparse)
    y2 = net.sample(gb, feat=feat[:1, :8].detach().expand(B, -1)) * (rast[..., 3:] > 0)
    assert float((y1 - y2).detach().abs().max()) < 1e-5
    net = cls(cin=3, cout=4, num_layers=3, nf=32)
    assert float((R._sample_field(net, gb, None, sparse) - net.sample(gb) * (rast[..., 3:] > 0)).detach().abs().max()) < 1e-6


def test_dmtet_geometry_torch_side_matches_reference_golden():
    """R1's PyTorch half (DMTetGeometry: get_sdf for every init mode, the BCE edge regulariser, the eikonal sample gradients,
    both regulariser values and their parameter gradients incl. the double backward, getAABB) against the reference class's
    own outputs (tests/golden/dmtet_geometry.npz).  The grid load / extraction (CUDA) is bypassed: only device-agnostic code runs."""
    g = golden("dmtet_geometry.npz")
    D = pkg("geometry.dmtet")
    orig = D.DMTetGeometry.load_tets

    def load_tets(self, grid_res=None, scale=None):
        self.verts = torch.from_numpy(g["verts"]) * self.grid_scale
        self.indices = torch.from_numpy(g["tets"])
        self.grid, self._all_edges = None, torch.from_numpy(g["all_edges"])

    D.DMTetGeometry.load_tets = load_tets
    try:
        geo = D.DMTetGeometry(6, 7.0, num_layers=5, hidden_size=32, embedder_freq=8, embed_concat_pts=True, init_sdf="ellipsoid",
                              jitter_grid=0.05, symmetrize=True)
    finally:
        D.DMTetGeometry.load_tets = orig
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
    assert set(sd) == set(geo.mlp.state_dict())                       # checkpoint-compatible parameter names
    geo.mlp.load_state_dict(sd)
    pts = torch.from_numpy(g["pts"])
    for mode in ("ellipsoid", "sphere", 0.25, None):
        geo.init_sdf = mode
        assert np.allclose(geo.get_sdf(pts).detach().numpy(), g["sdf_%s" % mode], atol=2e-6)
        assert np.allclose(geo.get_sdf().detach().numpy(), g["sdf_grid_%s" % mode], atol=2e-6)
    geo.init_sdf = "bogus"
    with pytest.raises(NotImplementedError):
        geo.get_sdf(pts)
    geo.init_sdf = "ellipsoid"
    geo.symmetrize = False
    assert np.allclose(geo.get_sdf(pts).detach().numpy(), g["sdf_nosym"], atol=2e-6)
    geo.symmetrize = True
    geo.current_sdf = geo.get_sdf()
    assert np.allclose(D.sdf_bce_reg_loss(geo.current_sdf, geo.all_edges).detach().numpy(), g["bce"], rtol=1e-5)
    geo.mesh_verts = torch.from_numpy(g["mesh_verts"])
    torch.manual_seed(43)
    grad = geo.get_sdf_gradient()
    assert grad.shape == (10000, 3) and np.allclose(grad.detach().numpy(), g["eikonal_grad"], atol=5e-5)
    torch.manual_seed(43)
    reg = geo.get_sdf_reg_loss()
    assert np.allclose(reg["sdf_bce_reg_loss"].detach().numpy(), g["reg_bce"], rtol=1e-5)
    assert np.allclose(reg["sdf_gradient_reg_loss"].detach().numpy(), g["reg_grad"], rtol=1e-4)
    (reg["sdf_bce_reg_loss"] + reg["sdf_gradient_reg_loss"]).backward()
    for k, p in geo.mlp.named_parameters():
        want = g["grad:" + k]
        assert np.abs(p.grad.numpy() - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), k
    with torch.no_grad():                                              # validation mode: zeros instead of an autograd error (:277-278)
        assert float(geo.get_sdf_gradient().abs().max()) == 0.0
    lo, hi = geo.getAABB()
    assert np.array_equal(torch.stack([lo, hi]).numpy(), g["aabb"])
    # the per-grid UV atlas of map_uv (dmtet.py:69-84), built once: 4 N^2 rows, N = ceil(sqrt(T))
    m = golden("mt_ellipsoid_12.npz")
    uv = D.DMTet(device="cpu").uv_table(6 * 12 ** 3, torch.device("cpu")).numpy()
    assert tuple(uv.shape) == tuple(m["uvs_shape"]) and np.allclose(uv[:64], m["uvs_head"], atol=1e-7)


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference tree only exists in the build container")
def test_mesh_utilities_match_reference_side_by_side():
    """The device-agnostic helpers of model/render/mesh.py run from the reference's own file next to this package's on the same
    CPU tensors: aabb, compute_edges (+inverse), unit_size, center_by_reference, compute_tangents (mesh.py:190-350).
    compute_edge_to_face_mapping hard-codes .cuda() upstream (:237-239), so its table is checked against its definition."""
    from oracle import reference_loader
    ref = reference_loader.load().mesh
    ours = pkg("render.mesh")
    rng = np.random.RandomState(12)
    B, V, F, Vt = 3, 40, 70, 55
    v_pos = torch.from_numpy(rng.randn(B, V, 3).astype(np.float32))
    v_nrm = torch.nn.functional.normalize(torch.from_numpy(rng.randn(B, V, 3).astype(np.float32)), dim=-1)
    v_tex = torch.from_numpy(rng.rand(B, Vt, 2).astype(np.float32))
    t_pos = torch.from_numpy(np.stack([rng.permutation(V)[:3] for _ in range(F)]).astype(np.int64))[None]
    t_pos[0, :V // 3 + 1] = torch.arange(3 * (V // 3 + 1)).remainder(V).view(-1, 3)      # every vertex is used
    t_tex = torch.from_numpy(rng.randint(0, Vt, (1, F, 3)).astype(np.int64))
    mk = lambda mod: mod.Mesh(v_pos.clone(), t_pos, v_nrm=v_nrm.clone(), t_nrm_idx=t_pos, v_tex=v_tex.clone(), t_tex_idx=t_tex)
    a, b = mk(ours), mk(ref)
    for x, y in zip(ours.aabb(a), ref.aabb(b)):
        assert torch.equal(x, y)
    assert torch.equal(ours.compute_edges(t_pos), ref.compute_edges(t_pos))
    for x, y in zip(ours.compute_edges(t_pos, return_inverse=True), ref.compute_edges(t_pos, return_inverse=True)):
        assert torch.equal(x, y)
    assert torch.equal(ours.unit_size(a).v_pos, ref.unit_size(b).v_pos)
    box = (torch.tensor([-1.0, -2.0, -0.5]), torch.tensor([2.0, 1.0, 0.5]))
    assert torch.equal(ours.center_by_reference(a, box, 1.7).v_pos, ref.center_by_reference(b, box, 1.7).v_pos)
    ta, tb = ours.compute_tangents(a), ref.compute_tangents(b)
    assert torch.allclose(ta.v_tng, tb.v_tng, atol=1e-6) and torch.equal(ta.t_tng_idx, tb.t_tng_idx)
    assert torch.allclose(a.v_tng, tb.v_tng, atol=1e-6)                 # and the lazily materialised attribute is the same thing
    # edge -> (triangle seeing it ascending, triangle seeing it descending); last writer wins exactly as upstream's indexing
    table = ours.compute_edge_to_face_mapping(t_pos)
    edges = ours.compute_edges(t_pos)
    want = np.zeros((edges.shape[0], 2), np.int64)
    lut = {tuple(e): i for i, e in enumerate(edges.tolist())}
    for f, tri in enumerate(t_pos[0].tolist()):
        for k in range(3):
            p, q = tri[k], tri[(k + 1) % 3]
            want[lut[(min(p, q), max(p, q))], int(p > q)] = f
    multi = np.zeros_like(want)
    for f, tri in enumerate(t_pos[0].tolist()):
        for k in range(3):
            p, q = tri[k], tri[(k + 1) % 3]
            multi[lut[(min(p, q), max(p, q))], int(p > q)] += 1
    once = multi <= 1                                                    # slots written by several triangles are order-dependent upstream too
    assert np.array_equal(table.numpy()[once], want[once]) and once.mean() > 0.9


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference tree only exists in the build container")
def test_skinning_helpers_match_reference_side_by_side():
    """The host helpers of model/geometry/skinning.py next to the reference's own functions on the same inputs: kinematic-chain
    builders (:25-46), children_to_parents (:273-282), _joints_to_bones (:8-13), Euler conventions and argument errors (:285-340)."""
    import copy
    import itertools
    from oracle import reference_loader
    ref = reference_loader.load().skinning
    ours = pkg("geometry.skinning")
    for n, start in ((8, 0), (7, 3), (2, 0), (1, 5)):
        assert ours.build_kinematic_chain(n, start) == ref.build_kinematic_chain(n, start)
    for n_body, n_leg, attach in ((8, 3, True), (8, 3, False), (4, 2, True), (6, 1, True)):
        half = n_body // 2
        chains = []
        for mod in (ours, ref):
            kc = mod.build_kinematic_chain(half, half)[1] + mod.build_kinematic_chain(half, 0)[1]
            legs = []
            for i in range(4):
                b2j, leg, dep = mod.build_kinematic_chain(n_leg, n_body + i * n_leg)
                kc = mod.update_body_kinematic_chain(copy.deepcopy(kc), leg, i % n_body, dep, attach_legs_to_body=attach)
                legs.append((b2j, leg, dep))
            chains.append((kc, legs, mod.children_to_parents(kc)))
        assert chains[0] == chains[1]
    joints = torch.arange(2 * 2 * 9 * 3, dtype=torch.float32).view(2, 2, 9, 3)
    idx = [(0, 1), (1, 2), (4, 3), (8, 0)]
    assert torch.equal(ours._joints_to_bones(joints, idx), ref._joints_to_bones(joints, idx))
    ang = torch.from_numpy(np.random.RandomState(4).uniform(-3, 3, (5, 2, 3)).astype(np.float32))
    for conv in ("".join(p) for p in itertools.product("XYZ", repeat=3)):
        try:
            want = ref.euler_angles_to_matrix(ang, conv)
        except ValueError as e:
            with pytest.raises(ValueError) as got:
                ours.euler_angles_to_matrix(ang, conv)
            assert str(got.value) == str(e)
            continue
        assert torch.allclose(ours.euler_angles_to_matrix(ang, conv), want, atol=1e-6), conv
    for bad_angles, conv in ((ang[..., :2], "XYZ"), (ang, "XY"), (ang, "XYW")):
        with pytest.raises(ValueError) as e1:
            ref.euler_angles_to_matrix(bad_angles, conv)
        with pytest.raises(ValueError) as e2:
            ours.euler_angles_to_matrix(bad_angles, conv)
        assert str(e1.value) == str(e2.value)
    for axis in "XYZ":
        assert torch.allclose(ours._axis_angle_rotation(axis, ang[..., 0]), ref._axis_angle_rotation(axis, ang[..., 0]), atol=1e-7)


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference tree only exists in the build container")
def test_network_twins_match_reference_classes():
    """The standalone twins in networks.py (used by bench.py --mlps and the tests when the reference tree is absent) are the
    reference's modules: same state-dict names and shapes, and with the same weights the same outputs (MLPs.py:9-101,
    HarmonicEmbedding.py) - so an M1b number measured on the twins is a number for the reference's field networks."""
    from oracle import reference_loader
    ref = reference_loader.load().mlps
    ours = pkg("networks")
    torch.manual_seed(9)
    x = torch.randn(4, 7, 3)
    mm = torch.tensor([[0.0, 1.0], [-1.0, 1.0], [0.2, 0.4]])
    cases = [("MLP", dict(cin=16, cout=4, num_layers=3, nf=32, activation="sigmoid"), torch.randn(5, 16), None),
             ("MLP", dict(cin=16, cout=4, num_layers=1), torch.randn(5, 16), None),
             ("CoordMLP", dict(cin=3, cout=3, num_layers=5, nf=32, activation="sigmoid", min_max=mm, n_harmonic_functions=6, embedder_scalar=0.8,
                               extra_feat_dim=8, symmetrize=True), x, torch.randn(4, 8)),
             ("CoordMLP", dict(cin=3, cout=1, num_layers=5, nf=32, n_harmonic_functions=8, embedder_scalar=0.8, embed_concat_pts=False), x, None),
             ("CoordMLP", dict(cin=3, cout=2, num_layers=2, nf=16, n_harmonic_functions=0, in_layer_relu=True, activation="tanh"), x, None)]
    for name, kw, inp, feat in cases:
        a, b = getattr(ours, name)(**kw), getattr(ref, name)(**kw)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb) and all(sa[k].shape == sb[k].shape for k in sa), name
        a.load_state_dict(sb)
        ya = a(inp) if feat is None else a(inp, feat=feat)
        yb = b(inp) if feat is None else b(inp, feat=feat)
        assert torch.allclose(ya, yb, atol=1e-6), (name, kw)
    with pytest.raises(NotImplementedError):
        ours.CoordMLP_Mod(3, 1, 5)


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference tree only exists in the build container")
def test_synthetic_projection_is_the_references():
    """The synthetic cameras use the reference's projection (render/util.py:189-194, negated y row: image row 0 is clip y = -1)
    and its safe_normalize / dot helpers agree with this package's mesh helpers."""
    from oracle import reference_loader
    ru = reference_loader.load().rutil
    syn = pkg("synthetic")
    for fov, asp, n, f in ((25 / 180 * np.pi, 1.0, 0.1, 1000.0), (0.7854, 1.5, 0.5, 50.0)):
        assert np.allclose(syn.perspective(fov, asp, n, f), ru.perspective(fov, asp, n, f).numpy(), rtol=1e-7, atol=0)
    m = pkg("render.mesh")
    x = torch.randn(5, 3)
    x[0] = 0
    assert torch.equal(m._safe_normalize(x), ru.safe_normalize(x)) and torch.equal(m._dot(x, x.flip(0)), ru.dot(x, x.flip(0)))
