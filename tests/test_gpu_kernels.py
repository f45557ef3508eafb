"""GPU parity: every kernel of libb2a.so (called through the C-ABI wrappers in 3danimals_b200.ops) against the CPU
oracle on the same seeded inputs.  Bar: bit-exact for index buffers (faces, uv indices, triangle ids, adjacency);
<= 1e-4 relative (of the tensor's max magnitude) for fp32 buffers and gradients."""
import numpy as np
import pytest
import torch

from conftest import assert_elem, golden, golden_files, pkg, rel_err
from oracle import geometry_np as gnp
from oracle import raster as R
from oracle import torch_ref as T

pytestmark = pytest.mark.gpu
TOL = 1e-4

syn = pkg("synthetic")


def _ops():
    return pkg("ops")


def dev(x, cuda):
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


# ----------------------------------------------------------------------------------------------------------------
# marching tets
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_files("mt_"))
def test_marching_tets_golden(cuda, name):
    """Against the reference's own DMTet output (committed fixture)."""
    ops = _ops()
    g = golden(name)
    v, t = syn.kuhn_tet_grid(int(g["res"]))
    v = v * np.float32(7.0)
    grid = ops.TetGrid(dev(t, cuda), v.shape[0])
    sdf = dev(g["sdf"], cuda)[:, None].requires_grad_(True)
    verts, faces, uv_idx, faces32, _ = ops.marching_tets(dev(v, cuda), sdf, grid)
    assert np.array_equal(faces.cpu().numpy(), g["faces"])
    assert np.array_equal(faces32.cpu().numpy(), g["faces"])
    assert np.array_equal(uv_idx.cpu().numpy(), g["uv_idx"])
    assert rel_err(verts.detach().cpu().numpy(), g["verts"]) < 1e-6
    (verts * dev(g["d_verts"], cuda)).sum().backward()
    assert rel_err(sdf.grad.cpu().numpy().reshape(-1), g["d_sdf"]) < TOL


@pytest.mark.parametrize("res", [16, 32, 64])
@pytest.mark.parametrize("kind", ["ellipsoid", "noisy_sphere", "horse"])
def test_marching_tets_oracle(cuda, res, kind):
    ops = _ops()
    v, t = syn.kuhn_tet_grid(res)
    v = v * np.float32(7.0)
    sdf = {"ellipsoid": syn.sdf_ellipsoid(v), "noisy_sphere": syn.sdf_noisy_sphere(v, 1.75, 0.05, 1),
           "horse": syn.sdf_horse(v, 0.01, 2)}[kind]
    o = gnp.marching_tets(v, sdf, t, with_uvs=False)
    grid = ops.TetGrid(dev(t, cuda), v.shape[0])
    assert np.array_equal(grid.all_edges().cpu().numpy(), gnp.unique_sorted_edges(t))
    sdf_d = dev(sdf, cuda).requires_grad_(True)
    pos_d = dev(v, cuda).requires_grad_(True)
    verts, faces, uv_idx, faces32, vert_edge = ops.marching_tets(pos_d, sdf_d, grid)
    assert np.array_equal(faces.cpu().numpy(), o["faces"])            # bit-exact
    assert np.array_equal(uv_idx.cpu().numpy(), o["uv_idx"])
    assert np.array_equal(vert_edge.cpu().numpy(), o["interp_v"])
    assert np.array_equal(verts.detach().cpu().numpy(), o["verts"])   # same unfused fp32 op order -> identical bits
    gv = np.random.RandomState(3).randn(*o["verts"].shape).astype(np.float32)
    (verts * dev(gv, cuda)).sum().backward()
    d_sdf, d_pos = gnp.lerp_vertices_bwd(v, sdf, o["interp_v"], gv)
    assert rel_err(sdf_d.grad.cpu().numpy(), d_sdf) < TOL
    assert rel_err(pos_d.grad.cpu().numpy(), d_pos) < TOL


def test_marching_tets_empty(cuda):
    ops = _ops()
    v, t = syn.kuhn_tet_grid(8)
    grid = ops.TetGrid(dev(t, cuda), v.shape[0])
    verts, faces, uv_idx, faces32, _ = ops.marching_tets(dev(v, cuda), torch.full((v.shape[0],), -1.0, device=cuda), grid)
    assert verts.shape == (0, 3) and faces.shape == (0, 3) and uv_idx.shape == (0, 3)


# ----------------------------------------------------------------------------------------------------------------
# LBS
# ----------------------------------------------------------------------------------------------------------------
def _chain(g):
    return [(int(b), [int(x) for x in str(d).split(",") if x != ""]) for b, d in zip(g["chain_ids"], g["chain_dep"])]


@pytest.mark.parametrize("name", golden_files("skin_"))
def test_skinning_golden(cuda, name):
    """Drop-in skinning() against the reference's own skinning output + autograd gradients (fixture)."""
    sk = pkg("geometry.skinning")
    g = golden(name)
    chain = _chain(g)
    vp = dev(g["verts"], cuda)[None, None].clone().requires_grad_(True)
    ang = dev(g["angles"], cuda).requires_grad_(True)
    out, aux = sk.skinning(vp, dev(g["bones"], cuda), chain, ang, output_posed_bones=True, temperature=0.05)
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < TOL
    assert rel_err(aux["posed_bones"].detach().cpu().numpy(), g["posed_bones"]) < TOL
    assert np.abs(aux["vertices_to_bones"].cpu().numpy() - g["weights"]).max() < TOL
    # per ELEMENT: |a - b| <= 1e-4 |b| + 1e-6 max|b| (the worst element is reported on failure)
    assert_elem(out.detach().cpu().numpy(), g["out"], what="posed vertices")
    assert_elem(aux["posed_bones"].detach().cpu().numpy(), g["posed_bones"], what="posed bones")
    ((out * dev(g["g_out"], cuda)).sum() + (aux["posed_bones"] * dev(g["g_posed"], cuda)).sum()).backward()
    assert rel_err(ang.grad.cpu().numpy(), g["d_angles"]) < TOL
    assert rel_err(vp.grad.cpu().numpy(), g["d_verts"]) < TOL


@pytest.mark.parametrize("name", golden_files("skin_"))
def test_estimate_bones_golden(cuda, name):
    sk = pkg("geometry.skinning")
    g = golden(name)
    verts = dev(g["verts"], cuda)[None, None]
    n_leg, mode = int(g["n_leg_bones"]), str(g["mode"])
    bones, chain, aux = sk.estimate_bones(verts, 8, n_legs=4, n_leg_bones=n_leg, body_bones_mode=mode)
    assert [(b, list(d)) for b, d in chain] == _chain(g)
    assert np.allclose(bones.cpu().numpy(), g["bones"], atol=1e-5)
    bones2 = sk.estimate_bones(verts * 1.01, 8, n_legs=4, n_leg_bones=n_leg, body_bones_mode=mode, compute_kinematic_chain=False, aux=aux)
    assert np.allclose(bones2.cpu().numpy(), g["bones_rescaled"], atol=1e-5)


@pytest.mark.parametrize("n_leg,mode", [(3, "z_minmax_y+"), (0, "z_minmax"), (2, "z_minmax")])
def test_estimate_bones_batched_oracle(cuda, n_leg, mode):
    """Fused kernel, B x F > 1 (per-instance deformation on, InstancePredictorBase.py:514-518): whole-batch x quantiles by
    radix select, per-(b,f) masked arg-min, attachment index fixed by the first (b,f) (skinning.py:190-192) vs the oracle."""
    sk = pkg("geometry.skinning")
    g = golden("skin_horse.npz")
    rng = np.random.RandomState(0)
    shapes = np.stack([g["verts"] * np.float32(1 + 0.05 * i) + rng.randn(*g["verts"].shape).astype(np.float32) * 0.01
                       for i in range(6)]).reshape(3, 2, -1, 3)
    ref_b, ref_chain, ref_aux = gnp.estimate_bones(shapes, 8, n_legs=4, n_leg_bones=n_leg, body_bones_mode=mode)
    bones, chain, aux = sk.estimate_bones(dev(shapes, cuda), 8, n_legs=4, n_leg_bones=n_leg, body_bones_mode=mode)
    assert [(b, list(d)) for b, d in chain] == [(b, list(d)) for b, d in ref_chain]
    assert np.allclose(bones.cpu().numpy(), ref_b, atol=1e-5)
    if n_leg:
        assert [l["body_bone_idx"] for l in aux["legs"]] == [l["body_bone_idx"] for l in ref_aux["legs"]]
    bones2 = sk.estimate_bones(dev(shapes, cuda) * 1.01, 8, n_legs=4, n_leg_bones=n_leg, body_bones_mode=mode, compute_kinematic_chain=False,
                               aux=aux)
    ref2 = gnp.estimate_bones(shapes * np.float32(1.01), 8, n_legs=4, n_leg_bones=n_leg, body_bones_mode=mode, compute_kinematic_chain=False,
                              aux=ref_aux)
    assert np.allclose(bones2.cpu().numpy(), ref2, atol=1e-5)
    # configured attachment joints (ponymation.yaml:114)
    b3, c3, a3 = sk.estimate_bones(dev(shapes, cuda), 8, n_legs=4, n_leg_bones=max(n_leg, 1), body_bones_mode=mode,
                                   legs_to_body_joint_indices=[2, 7, 7, 2])
    r3, rc3, ra3 = gnp.estimate_bones(shapes, 8, n_legs=4, n_leg_bones=max(n_leg, 1), body_bones_mode=mode, legs_to_body_joint_indices=[2, 7, 7, 2])
    assert np.allclose(b3.cpu().numpy(), r3, atol=1e-5) and [(b, list(d)) for b, d in c3] == [(b, list(d)) for b, d in rc3]


@pytest.mark.parametrize("tag", ["single", "batch"])
def test_estimate_bones_fauna_variant_golden(cuda, tag):
    """3D-Fauna's bone_y_threshold variant (config C3) on the device vs the reference's golden."""
    sk = pkg("geometry.skinning")
    g = golden("bones_fauna.npz")
    shape = dev(g[tag + "_shape"], cuda)
    bones, chain, aux = sk.estimate_bones(shape, 8, n_legs=4, n_leg_bones=3, body_bones_mode="z_minmax_y+", bone_y_threshold=0.4)
    ref_chain = [(int(b), [int(x) for x in str(d).split(",") if x != ""]) for b, d in zip(g[tag + "_chain_ids"], g[tag + "_chain_dep"])]
    assert [(int(b), [int(x) for x in d]) for b, d in chain] == ref_chain
    assert np.allclose(bones.cpu().numpy(), g[tag + "_bones"], atol=1e-5)
    bones2 = sk.estimate_bones(shape * 1.01, 8, n_legs=4, n_leg_bones=3, body_bones_mode="z_minmax_y+", compute_kinematic_chain=False,
                               aux=aux, bone_y_threshold=0.4)
    assert np.allclose(bones2.cpu().numpy(), g[tag + "_bones_rescaled"], atol=1e-5)


def test_estimate_bones_fauna_quantiles_exact_and_batched_oracle(cuda):
    """The Fauna variant's seven radix-select quantiles (one over everything, six over the data-dependent subset y < y_thr)
    equal torch.quantile on the same device bits (stats hook), and random batched shapes agree with the numpy oracle."""
    ops, sk, lib = _ops(), pkg("geometry.skinning"), pkg("_lib")
    rng = np.random.RandomState(5)
    for n_inst, V in ((1, 1500), (3, 4099)):
        x = dev(rng.randn(n_inst, 1, V, 3).astype(np.float32) * 2, cuda)
        x[0, 0, :9, 1] = x[0, 0, 9, 1]          # ties around / inside the subset
        ws = torch.empty(ops._size(lib.lib().b2a_estimate_bones_workspace_bytes), dtype=torch.uint8, device=cuda)
        bones = torch.empty(n_inst, 20, 2, 3, device=cuda)
        stats = torch.zeros(n_inst, 8, device=cuda)
        ops._call("b2a_estimate_bones", (x.data_ptr(), n_inst, V, 8, 3, 1, 0.4, -1, -1, -1, -1, ws.data_ptr(), ws.numel(), bones.data_ptr(), None,
                                         stats.data_ptr(), ops._stream()), launches=7)
        xs, ys, zs = x[..., 0], x[..., 1], x[..., 2]
        flags = ys < ys.quantile(0.4)
        xm = (xs[flags].quantile(0.95) - xs[flags].quantile(0.05)) * 0.2
        assert torch.all(stats[:, 0] == xm), (stats[:, 0], xm)
        assert torch.all(stats[:, 6] == xs[flags].quantile(0.5)) and torch.all(stats[:, 7] == zs[flags].quantile(0.5))
    g = golden("skin_horse.npz")
    shapes = np.stack([g["verts"] * np.float32(1 + 0.05 * i) + rng.randn(*g["verts"].shape).astype(np.float32) * 0.01
                       for i in range(4)]).reshape(2, 2, -1, 3)
    ref_b, ref_chain, ref_aux = gnp.estimate_bones(shapes, 8, n_legs=4, n_leg_bones=3, body_bones_mode="z_minmax_y+", bone_y_threshold=0.4)
    bones, chain, aux = sk.estimate_bones(dev(shapes, cuda), 8, n_legs=4, n_leg_bones=3, body_bones_mode="z_minmax_y+", bone_y_threshold=0.4)
    assert [(int(b), [int(v) for v in d]) for b, d in chain] == [(int(b), [int(v) for v in d]) for b, d in ref_chain]
    assert np.allclose(bones.cpu().numpy(), ref_b, atol=1e-5)
    assert ops.stats.calls.get("b2a_estimate_bones", 0) > 0


def test_estimate_bones_quantile_exact(cuda):
    """The radix-select quantiles equal torch.quantile on the same device bits (x_margin exposed through the stats hook)."""
    ops = _ops()
    lib = pkg("_lib")
    rng = np.random.RandomState(3)
    for n_inst, V in ((1, 1000), (4, 2357), (2, 50001)):
        x = dev(rng.randn(n_inst, 1, V, 3).astype(np.float32) * 3, cuda)
        x[0, 0, :7, 0] = x[0, 0, 7, 0]          # ties
        ws = torch.empty(ops._size(lib.lib().b2a_estimate_bones_workspace_bytes), dtype=torch.uint8, device=cuda)
        bones = torch.empty(n_inst, 20, 2, 3, device=cuda)
        stats = torch.zeros(n_inst, 8, device=cuda)
        ops._call("b2a_estimate_bones", (x.data_ptr(), n_inst, V, 8, 3, 1, 0.0, -1, -1, -1, -1, ws.data_ptr(), ws.numel(), bones.data_ptr(), None,
                                         stats.data_ptr(), ops._stream()), launches=4)
        xs = x[..., 0]
        ref = (xs.quantile(0.95) - xs.quantile(0.05)) * 0.2
        assert torch.all(stats[:, 0] == ref), (stats[:, 0], ref)
        assert torch.allclose(stats[:, 1:4], x.mean(2)[:, 0], atol=1e-6)


@pytest.mark.parametrize("Bv,Bb", [(1, 1), (3, 1), (3, 3)])
def test_lbs_batched_oracle(cuda, Bv, Bb):
    ops = _ops()
    g = golden("skin_horse.npz")
    chain = _chain(g)
    rng = np.random.RandomState(4)
    B, K = 3, g["bones"].shape[2]
    verts = np.stack([g["verts"] * np.float32(1 + 0.02 * i) for i in range(Bv)])[:, None]          # [Bv,1,V,3]
    bones = np.concatenate([g["bones"] * np.float32(1 + 0.02 * i) for i in range(Bb)], 0)          # [Bb,1,K,2,3]
    ang = rng.uniform(-0.4, 0.4, (B, 1, K, 3)).astype(np.float32)
    vt = torch.from_numpy(verts).requires_grad_(True)
    at = torch.from_numpy(ang).requires_grad_(True)
    ref, raux = T.skinning(vt, torch.from_numpy(bones), chain, at, temperature=0.05)
    go = rng.randn(*ref.shape).astype(np.float32)
    gp = rng.randn(B, 1, K, 2, 3).astype(np.float32)
    ((ref * torch.from_numpy(go)).sum() + (raux["posed_bones"] * torch.from_numpy(gp)).sum()).backward()
    cp, ci = ops.chain_tables(chain, K, cuda)
    vd = dev(verts[:, 0], cuda).requires_grad_(True)
    ad = dev(ang[:, 0], cuda).requires_grad_(True)
    out, posed, w = ops.lbs(vd, dev(bones[:, 0], cuda), ad, cp, ci, 0.05, want_weights=True)
    assert rel_err(out.detach().cpu().numpy(), ref.detach().numpy()[:, 0]) < TOL
    assert rel_err(posed.detach().cpu().numpy(), raux["posed_bones"].detach().numpy()[:, 0]) < TOL
    assert np.abs(w.cpu().numpy() - raux["vertices_to_bones"].detach().numpy()[:, :, 0]).max() < TOL
    ((out * dev(go[:, 0], cuda)).sum() + (posed * dev(gp[:, 0], cuda)).sum()).backward()
    assert rel_err(ad.grad.cpu().numpy(), at.grad.numpy()[:, 0]) < TOL
    assert rel_err(vd.grad.cpu().numpy(), vt.grad.numpy()[:, 0]) < TOL


@pytest.mark.parametrize("res,dtype", [(12, torch.int64), (20, torch.int32), (33, torch.int64)])
def test_static_grid_tables(cuda, res, dtype):
    """b2a_mt_build_edges / b2a_mt_emit_edges / b2a_mt_build_tile_words (the library's replacement of generate_edges, dmtet.py:283-288,
    + the tile skip table) bit-equal to the torch.unique formulation; all_edges equals the reference's own edge list."""
    from oracle import torch_ops_geometry as G
    import ctypes
    ops, lib = _ops(), pkg("_lib")
    v, t = syn.kuhn_tet_grid(res)
    tets = dev(t, cuda).to(dtype)
    grid = ops.TetGrid(tets, v.shape[0])
    tt, tw = ctypes.c_int(0), ctypes.c_int(0)
    lib.check(lib.lib().b2a_mt_tile_shape(ctypes.byref(tt), ctypes.byref(tw)))
    start, edge_b, table = G.static_tables(tets, v.shape[0], tt.value, tw.value)
    assert grid.E == edge_b.numel() and torch.equal(grid.edge_start, start) and torch.equal(grid.edge_b, edge_b)
    assert torch.equal(grid.tile_words, table)
    # the reference's formulation verbatim (dmtet.py:283-288)
    be = torch.tensor([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], device=cuda)
    all_edges = tets.long()[:, be].reshape(-1, 2)
    all_edges = torch.unique(torch.sort(all_edges, dim=1)[0], dim=0)
    assert torch.equal(grid.all_edges(), all_edges)


# ----------------------------------------------------------------------------------------------------------------
# normals, clip transform
# ----------------------------------------------------------------------------------------------------------------
def _mesh(res=20, batch=2, seed=0):
    v, t = syn.kuhn_tet_grid(res)
    v = v * np.float32(7.0)
    o = gnp.marching_tets(v, syn.sdf_horse(v, 0.0, 0), t, with_uvs=False)
    rng = np.random.RandomState(seed)
    verts = np.stack([o["verts"] + rng.randn(*o["verts"].shape).astype(np.float32) * 0.003 for _ in range(batch)])
    return verts.astype(np.float32), o["faces"].astype(np.int32), o["verts"]


def test_vertex_normals_oracle(cuda):
    ops = _ops()
    verts, faces, _ = _mesh()
    vt = torch.from_numpy(verts).requires_grad_(True)
    ref = T.auto_normals(vt, torch.from_numpy(faces).long())       # the reference's torch ops (mesh.py:276-304)
    g = np.random.RandomState(1).randn(*verts.shape).astype(np.float32)
    (ref * torch.from_numpy(g)).sum().backward()
    vd = dev(verts, cuda).requires_grad_(True)
    nrm = ops.vertex_normals(vd, dev(faces, cuda))
    assert rel_err(nrm.detach().cpu().numpy(), ref.detach().numpy()) < TOL
    (nrm * dev(g, cuda)).sum().backward()
    assert rel_err(vd.grad.cpu().numpy(), vt.grad.numpy()) < TOL
    # degenerate input: isolated vertex gets the (0,0,1) fallback
    v2 = np.concatenate([verts, np.zeros((2, 1, 3), np.float32)], 1)
    n2 = ops.vertex_normals(dev(v2, cuda), dev(faces, cuda)).cpu().numpy()
    assert np.array_equal(n2[:, -1], np.array([[0, 0, 1], [0, 0, 1]], np.float32))


@pytest.mark.parametrize("case", ["mesh", "degenerate"])
def test_vertex_normals_golden(cuda, case):
    """R3 against the reference's own `auto_normals` (committed fixture tests/golden/normals.npz, generated by running
    model/render/mesh.py:276-304 unmodified): values, gradient, degenerate-vertex fallback."""
    ops = _ops()
    g = golden("normals.npz")
    v = dev(g[case + "_v_pos"], cuda).requires_grad_(True)
    nrm = ops.vertex_normals(v, dev(g[case + "_faces"], cuda))
    assert rel_err(nrm.detach().cpu().numpy(), g[case + "_v_nrm"]) < TOL
    assert_elem(nrm.detach().cpu().numpy(), g[case + "_v_nrm"], floor=1e-5, what="vertex normals")
    (nrm * dev(g[case + "_g"], cuda)).sum().backward()
    assert rel_err(v.grad.cpu().numpy(), g[case + "_d_v_pos"]) < TOL
    if case == "degenerate":
        assert np.array_equal(nrm.detach().cpu().numpy()[0, 3:], np.array([[0, 0, 1]] * 3, np.float32))


@pytest.mark.parametrize("Bp", [1, 3])
def test_xfm_points_oracle(cuda, Bp):
    ops = _ops()
    rng = np.random.RandomState(2)
    pts = rng.randn(Bp, 777, 3).astype(np.float32)
    mtx = rng.randn(3, 4, 4).astype(np.float32)
    pt = torch.from_numpy(pts).requires_grad_(True)
    mt = torch.from_numpy(mtx).requires_grad_(True)
    ref = T.xfm_points(pt, mt)
    g = rng.randn(3, 777, 4).astype(np.float32)
    (ref * torch.from_numpy(g)).sum().backward()
    pd, md = dev(pts, cuda).requires_grad_(True), dev(mtx, cuda).requires_grad_(True)
    out = ops.xfm_points(pd, md)
    assert np.array_equal(out.detach().cpu().numpy(), R.xfm_points(pts, mtx))   # same op order as the C oracle
    assert rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < TOL
    (out * dev(g, cuda)).sum().backward()
    assert rel_err(pd.grad.cpu().numpy(), pt.grad.numpy()) < TOL
    assert rel_err(md.grad.cpu().numpy(), mt.grad.numpy()) < TOL


def test_marching_tets_res256_capacity(cuda):
    """BASELINE configs[4] grid size (DMTet res 256: 17 M grid vertices, 100 M tets, 1.6 GB of tet indices): the extraction
    runs, agrees with a second run bit for bit, and yields a closed, consistently oriented surface of the expected size
    (vertex count scales ~4x per resolution doubling)."""
    ops = _ops()
    v, t = syn.kuhn_tet_grid_torch(256, cuda)
    v = v * 7.0
    q = v.clone()
    q[:, 2] = q[:, 2] / 2
    sdf = (1.05 - q.norm(dim=-1))[:, None].contiguous()          # the reference's ellipsoid init (dmtet.py:246-250)
    grid = ops.TetGrid(t, v.shape[0])
    del t, q
    assert grid.T == 6 * 256 ** 3 and grid.Vg == 257 ** 3
    assert int((grid.tile_words[:, 0] < 0).sum()) == 0          # every tile of the lattice fits the skip table
    verts, faces, uv_idx, faces32, vert_edge = ops.marching_tets(v, sdf, grid)
    verts2, faces2, _, _, _ = ops.marching_tets(v, sdf, grid)
    assert torch.equal(verts, verts2) and torch.equal(faces, faces2)
    V, F = verts.shape[0], faces.shape[0]
    assert 100_000 < V < 200_000 and F == 2 * V - 4               # genus-0 closed triangle mesh: F = 2V - 4
    e = torch.cat([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
    assert torch.unique(e[:, 0] * V + e[:, 1]).numel() == e.shape[0]                       # consistently oriented
    _, cnt = torch.unique(torch.minimum(e[:, 0], e[:, 1]) * V + torch.maximum(e[:, 0], e[:, 1]), return_counts=True)
    assert bool((cnt == 2).all())                                                          # closed 2-manifold
    r = (verts * torch.tensor([1.0, 1.0, 0.5], device=cuda)).norm(dim=-1)
    assert float((r - 1.05).abs().max()) < 1e-3                                            # vertices lie on the iso-surface
    assert torch.equal(uv_idx[:, 0] % 4, torch.zeros_like(uv_idx[:, 0]))


# ----------------------------------------------------------------------------------------------------------------
# rasterize / interpolate / antialias
# ----------------------------------------------------------------------------------------------------------------
def _scene(res=24, batch=3, image=96, seed=0):
    verts, faces, prior = _mesh(res, batch, seed)
    mvp, w2c, campos = syn.cameras(batch, seed=seed + 3)
    clip = R.xfm_points(verts, mvp)
    return verts, faces, prior, mvp, w2c, campos, clip


from raster_scenes import ANALYTIC  # noqa: E402


@pytest.mark.parametrize("name", sorted(ANALYTIC))
@pytest.mark.parametrize("res", [(16, 16), (37, 53)])
def test_rasterize_analytic(cuda, name, res):
    ops = _ops()
    pos, tri = ANALYTIC[name]
    ref = R.rasterize(pos, tri, res)
    out = ops.rasterize(dev(pos, cuda), dev(tri, cuda), res).cpu().numpy()
    assert np.array_equal(out[..., 3], ref[..., 3])            # bit-exact triangle ids
    assert np.array_equal(out, ref)                            # and identical barycentrics / depth bits


def test_rasterize_edge_and_depth_ties(cuda):
    """Pixel centres exactly on edges / on a shared diagonal / equal depth: the kernel breaks every tie like the restatement
    (inclusive edges, lowest triangle index) - tests/test_oracle_raster.py documents where that differs from nvdiffrast-GL."""
    from test_oracle_raster import edge_tie_scene
    ops = _ops()
    for res in (8, 16, 64):
        pos, tri = edge_tie_scene(res)
        ref = R.rasterize(pos, tri, (res, res))
        out = ops.rasterize(dev(pos, cuda), dev(tri, cuda), (res, res)).cpu().numpy()
        assert np.array_equal(out, ref)
        assert (out[0, ..., 3] > 0).sum() == (res - 2) ** 2 and not np.isin(out[0, ..., 3], [3.0, 4.0]).any()


@pytest.mark.parametrize("image", [64, 256])
def test_rasterize_mesh_oracle(cuda, image):
    ops = _ops()
    verts, faces, prior, mvp, w2c, campos, clip = _scene(image=image)
    ref = R.rasterize(clip, faces, (image, image))
    cd = dev(clip, cuda).requires_grad_(True)
    out = ops.rasterize(cd, dev(faces, cuda), (image, image))
    assert (ref[..., 3] > 0).mean() > 0.05
    assert np.array_equal(out.detach().cpu().numpy()[..., 3], ref[..., 3])
    assert np.array_equal(out.detach().cpu().numpy(), ref)
    g = np.random.RandomState(0).randn(*ref.shape).astype(np.float32)
    (out * dev(g, cuda)).sum().backward()
    assert rel_err(cd.grad.cpu().numpy(), R.rasterize_bwd(clip, faces, ref, g)) < TOL


@pytest.mark.parametrize("Ba,C", [(3, 3), (1, 3), (3, 16), (3, 1)])
def test_interpolate_oracle(cuda, Ba, C):
    ops = _ops()
    verts, faces, prior, mvp, w2c, campos, clip = _scene()
    rast = R.rasterize(clip, faces, (96, 96))
    rng = np.random.RandomState(5)
    attr = rng.randn(Ba, verts.shape[1], C).astype(np.float32)
    ref = R.interpolate(attr, rast, faces)
    ad = dev(attr, cuda).requires_grad_(True)
    rd = dev(rast, cuda).requires_grad_(True)
    out = ops.interpolate(ad, rd, dev(faces, cuda))
    assert np.array_equal(out.detach().cpu().numpy(), ref)
    g = rng.randn(*ref.shape).astype(np.float32)
    (out * dev(g, cuda)).sum().backward()
    da, dr = R.interpolate_bwd(attr, rast, faces, g)
    assert rel_err(ad.grad.cpu().numpy(), da) < TOL
    assert rel_err(rd.grad.cpu().numpy(), dr) < TOL


def test_edge_adjacency_oracle(cuda):
    ops = _ops()
    verts, faces, _ = _mesh(24)
    for tri, V in ((faces, verts.shape[1]), (ANALYTIC["shared_edge_fan"][1], 7), (ANALYTIC["overlapping_quads"][1], 8),
                   (np.array([[0, 1, 2], [2, 1, 0], [0, 1, 3], [1, 1, 2]], np.int32), 4)):   # duplicate + non-manifold + degenerate
        assert np.array_equal(ops.edge_adjacency(dev(tri, cuda), V).cpu().numpy(), R.edge_adjacency(tri, V))


@pytest.mark.parametrize("C", [4, 17, 2])
def test_antialias_oracle(cuda, C):
    ops = _ops()
    verts, faces, prior, mvp, w2c, campos, clip = _scene()
    rast = R.rasterize(clip, faces, (96, 96))
    rng = np.random.RandomState(6)
    color = rng.rand(3, 96, 96, C).astype(np.float32)
    color[..., -1] = (rast[..., 3] > 0)
    opp = R.edge_adjacency(faces, verts.shape[1])
    ref = R.antialias(color, rast, clip, faces, opp)
    assert np.abs(ref - color).max() > 0.05      # the case does blend something
    cd = dev(color, cuda).requires_grad_(True)
    pd = dev(clip, cuda).requires_grad_(True)
    out = ops.antialias(cd, dev(rast, cuda), pd, dev(faces, cuda))
    assert np.array_equal(out.detach().cpu().numpy(), ref)   # gather in the oracle's accumulation order -> identical bits
    g = rng.randn(*ref.shape).astype(np.float32)
    (out * dev(g, cuda)).sum().backward()
    dc, dp = R.antialias_bwd(color, rast, clip, faces, g, opp)
    assert rel_err(cd.grad.cpu().numpy(), dc) < TOL
    assert np.abs(dp).max() > 0
    assert rel_err(pd.grad.cpu().numpy(), dp) < TOL


@pytest.mark.parametrize("image,layout,prepared", [(96, "nchw", False), (96, "nchw", True), (128, "nchw", True), (128, "nhwc", True),
                                                   (100, "nchw", True)])   # prepared: stream + fix-up kernels; else generic
@pytest.mark.parametrize("C,keep,with_bg", [(4, 4, True), (17, 16, False), (2, 1, True), (3, 2, False)])
def test_composite_antialias_oracle(cuda, C, keep, with_bg, image, layout, prepared):
    """Fused lerp(bg,[color,1],id>0) + antialias + channel slice vs the unfused torch/oracle sequence (render.py:258-331),
    with the upstream gradient arriving NCHW-contiguous (a loss on the permuted view) or NHWC-contiguous."""
    ops = _ops()
    verts, faces, prior, mvp, w2c, campos, clip = _scene()
    S = image
    rast = R.rasterize(clip, faces, (S, S))
    rng = np.random.RandomState(7)
    color = rng.rand(3, S, S, C - 1).astype(np.float32)
    bg = rng.rand(3, S, S, C).astype(np.float32) if with_bg else None
    opp = R.edge_adjacency(faces, verts.shape[1])
    ct = torch.from_numpy(color).requires_grad_(True)
    pt = torch.from_numpy(clip).requires_grad_(True)
    alpha = torch.from_numpy((rast[..., 3:] > 0).astype(np.float32))
    bgt = torch.from_numpy(bg) if with_bg else torch.zeros(1, S, S, C)
    acc = torch.lerp(bgt.expand(3, -1, -1, -1), torch.cat((ct, torch.ones_like(ct[..., :1])), -1), alpha)
    ref = T.antialias(acc.contiguous(), torch.from_numpy(rast), pt, torch.from_numpy(faces), torch.from_numpy(opp))[..., :keep].permute(0, 3, 1, 2)
    g = rng.randn(*ref.shape).astype(np.float32)
    (ref * torch.from_numpy(g)).sum().backward()
    cd = dev(color, cuda).requires_grad_(True)
    pd = dev(clip, cuda).requires_grad_(True)
    aa_ctx = ops.antialias_prepare(dev(rast, cuda), pd.detach(), dev(faces, cuda), dev(opp, cuda)) if prepared else None
    assert (aa_ctx is not None) == (prepared and (S * S) % 32 == 0)
    out = ops.composite_antialias(cd, dev(bg, cuda) if with_bg else None, dev(rast, cuda), pd, dev(faces, cuda), dev(opp, cuda),
                                  True, keep, aa_ctx=aa_ctx).permute(0, 3, 1, 2)
    assert np.array_equal(out.detach().cpu().numpy(), ref.detach().numpy())
    gd = dev(g, cuda)
    if layout == "nhwc":
        gd = gd.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    out.backward(gd)
    assert rel_err(cd.grad.cpu().numpy(), ct.grad.numpy()) < TOL
    assert np.abs(pt.grad.numpy()).max() > 0
    assert rel_err(pd.grad.cpu().numpy(), pt.grad.numpy()) < TOL


@pytest.mark.parametrize("image,layout,keep_n", [(128, "nchw", 4), (128, "nhwc", 4), (96, "nchw", 3), (128, "strided", 4)])
def test_composite_antialias_pair_oracle(cuda, image, layout, keep_n):
    """The fused two-key launch (dino_pred 16+1 | shaded 3+1, b2a_antialias_pair_fwd/bwd) vs the oracle's two separate
    composite+antialias passes: images bit-exact, colour gradients and the SUMMED position gradient <= 1e-4.
    'strided': a wide gradient that is neither NCHW- nor NHWC-contiguous takes the two single-key launches."""
    ops = _ops()
    verts, faces, prior, mvp, w2c, campos, clip = _scene()
    S = image
    rast = R.rasterize(clip, faces, (S, S))
    rng = np.random.RandomState(11)
    col_w = rng.rand(3, S, S, 16).astype(np.float32)
    col_n = rng.rand(3, S, S, 3).astype(np.float32)
    bg_n = rng.rand(3, S, S, 4).astype(np.float32)
    opp = R.edge_adjacency(faces, verts.shape[1])
    pt = torch.from_numpy(clip).requires_grad_(True)
    alpha = torch.from_numpy((rast[..., 3:] > 0).astype(np.float32))
    refs, cts, gs = [], [], []
    for col, bg, C, keep in ((col_w, None, 17, 16), (col_n, bg_n, 4, keep_n)):
        ct = torch.from_numpy(col).requires_grad_(True)
        bgt = torch.from_numpy(bg) if bg is not None else torch.zeros(1, S, S, C)
        acc = torch.lerp(bgt.expand(3, -1, -1, -1), torch.cat((ct, torch.ones_like(ct[..., :1])), -1), alpha)
        ref = T.antialias(acc.contiguous(), torch.from_numpy(rast), pt, torch.from_numpy(faces), torch.from_numpy(opp))[..., :keep].permute(0, 3, 1, 2)
        g = rng.randn(*ref.shape).astype(np.float32)
        refs.append(ref); cts.append(ct); gs.append(g)
    torch.autograd.backward(refs, [torch.from_numpy(g) for g in gs])
    cw, cn = dev(col_w, cuda).requires_grad_(True), dev(col_n, cuda).requires_grad_(True)
    pd = dev(clip, cuda).requires_grad_(True)
    rd, fd, od = dev(rast, cuda), dev(faces, cuda), dev(opp, cuda)
    aa_ctx = ops.antialias_prepare(rd, pd.detach(), fd, od)
    assert aa_ctx is not None and ops.pair_supported(cw, cn, aa_ctx)
    ow, on = ops.composite_antialias_pair(cw, None, 16, cn, dev(bg_n, cuda), keep_n, rd, pd, fd, od, aa_ctx)
    ow, on = ow.permute(0, 3, 1, 2), on.permute(0, 3, 1, 2)
    assert np.array_equal(ow.detach().cpu().numpy(), refs[0].detach().numpy())
    assert np.array_equal(on.detach().cpu().numpy(), refs[1].detach().numpy())
    gw, gn = dev(gs[0], cuda), dev(gs[1], cuda)
    if layout == "nhwc":
        gw = gw.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        gn = gn.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    elif layout == "strided":
        gw = torch.stack([gw, gw], -1)[..., 0]       # x stride 2: no fused instantiation
    torch.autograd.backward([ow, on], [gw, gn])
    assert rel_err(cw.grad.cpu().numpy(), cts[0].grad.numpy()) < TOL
    assert rel_err(cn.grad.cpu().numpy(), cts[1].grad.numpy()) < TOL
    assert np.abs(pt.grad.numpy()).max() > 0
    assert rel_err(pd.grad.cpu().numpy(), pt.grad.numpy()) < TOL
    calls = ops.stats.all_calls()
    assert calls.get("b2a_antialias_pair_fwd", 0) >= 1


@pytest.mark.parametrize("C,keep,up,aa,with_bg", [(4, 4, 2, True, True), (4, 4, 4, True, False), (2, 1, 2, True, True), (4, 3, 1, False, False),
                                                  (4, 3, 2, False, True), (3, 2, 4, True, False)])
def test_composite_up_oracle(cuda, C, keep, up, aa, with_bg):
    """Low-resolution colour up-sampled inside the composite (+ antialias) kernel vs the reference's sequence: nearest
    upsample (render.py:217-219) -> lerp composite -> antialias -> channel slice; backward incl. the up x up block sum."""
    ops = _ops()
    verts, faces, prior, mvp, w2c, campos, clip = _scene()
    S = 128
    s = S // up
    rast = R.rasterize(clip, faces, (S, S))
    rng = np.random.RandomState(13)
    color = rng.rand(3, s, s, C - 1).astype(np.float32)
    bg = rng.rand(3, S, S, C).astype(np.float32) if with_bg else None
    opp = R.edge_adjacency(faces, verts.shape[1])
    ct = torch.from_numpy(color).requires_grad_(True)
    pt = torch.from_numpy(clip).requires_grad_(True)
    alpha = torch.from_numpy((rast[..., 3:] > 0).astype(np.float32))
    bgt = torch.from_numpy(bg) if with_bg else torch.zeros(1, S, S, C)
    cu = ct.repeat_interleave(up, dim=1).repeat_interleave(up, dim=2)
    acc = torch.lerp(bgt.expand(3, -1, -1, -1), torch.cat((cu, torch.ones_like(cu[..., :1])), -1), alpha)
    if aa:
        acc = T.antialias(acc.contiguous(), torch.from_numpy(rast), pt, torch.from_numpy(faces), torch.from_numpy(opp))
    ref = acc[..., :keep].permute(0, 3, 1, 2)
    g = rng.randn(*ref.shape).astype(np.float32)
    (ref * torch.from_numpy(g)).sum().backward()
    cd = dev(color, cuda).requires_grad_(True)
    pd = dev(clip, cuda).requires_grad_(True)
    aa_ctx = ops.antialias_prepare(dev(rast, cuda), pd.detach(), dev(faces, cuda), dev(opp, cuda))
    out = ops.composite_up(cd, dev(bg, cuda) if with_bg else None, pd, (S, S), up=up, antialias_edges=aa, keep=keep, aa_ctx=aa_ctx).permute(0, 3, 1, 2)
    assert np.array_equal(out.detach().cpu().numpy(), ref.detach().numpy())
    out.backward(dev(g, cuda))
    assert rel_err(cd.grad.cpu().numpy(), ct.grad.numpy()) < TOL
    if aa:
        assert np.abs(pt.grad.numpy()).max() > 0
        assert rel_err(pd.grad.cpu().numpy(), pt.grad.numpy()) < TOL
    else:
        assert pd.grad is None or float(pd.grad.abs().max()) == 0


@pytest.mark.parametrize("C,keep,up,aa,with_bg", [(4, 4, 2, True, True), (4, 4, 4, True, False), (2, 1, 4, True, True), (4, 3, 2, False, True),
                                                  (3, 2, 4, True, False)])
def test_composite_up_pool_oracle(cuda, C, keep, up, aa, with_bg):
    """The msaa resolve fused into the composite kernel (b2a_composite_up_pool_fwd/bwd) vs the reference's sequence: nearest upsample
    (render.py:217-219) -> lerp composite -> antialias -> channel slice -> avg_pool (render.py:322-323, util.avg_pool_nhwc); and
    bit-identical to the unfused library path (composite_up, then avg_pool2d)."""
    ops = _ops()
    verts, faces, prior, mvp, w2c, campos, clip = _scene()
    S = 128
    s = S // up
    rast = R.rasterize(clip, faces, (S, S))
    rng = np.random.RandomState(17)
    color = rng.rand(3, s, s, C - 1).astype(np.float32)
    bg = rng.rand(3, S, S, C).astype(np.float32) if with_bg else None
    opp = R.edge_adjacency(faces, verts.shape[1])
    ct = torch.from_numpy(color).requires_grad_(True)
    pt = torch.from_numpy(clip).requires_grad_(True)
    alpha = torch.from_numpy((rast[..., 3:] > 0).astype(np.float32))
    bgt = torch.from_numpy(bg) if with_bg else torch.zeros(1, S, S, C)
    cu = ct.repeat_interleave(up, dim=1).repeat_interleave(up, dim=2)
    acc = torch.lerp(bgt.expand(3, -1, -1, -1), torch.cat((cu, torch.ones_like(cu[..., :1])), -1), alpha)
    if aa:
        acc = T.antialias(acc.contiguous(), torch.from_numpy(rast), pt, torch.from_numpy(faces), torch.from_numpy(opp))
    ref = torch.nn.functional.avg_pool2d(acc[..., :keep].permute(0, 3, 1, 2), up)
    g = rng.randn(*ref.shape).astype(np.float32)
    (ref * torch.from_numpy(g)).sum().backward()
    res = []
    for pool in (True, False):
        cd = dev(color, cuda).requires_grad_(True)
        pd = dev(clip, cuda).requires_grad_(True)
        aa_ctx = ops.antialias_prepare(dev(rast, cuda), pd.detach(), dev(faces, cuda), dev(opp, cuda))
        ops.stats.reset()
        out = ops.composite_up(cd, dev(bg, cuda) if with_bg else None, pd, (S, S), up=up, antialias_edges=aa, keep=keep, aa_ctx=aa_ctx,
                               pool=pool).permute(0, 3, 1, 2)
        if not pool:
            out = torch.nn.functional.avg_pool2d(out, up)
        assert tuple(out.shape) == (3, keep, s, s)
        out.backward(dev(g, cuda))
        assert ("b2a_composite_up_pool_fwd" in ops.stats.calls) == pool and ("b2a_composite_up_pool_bwd" in ops.stats.calls) == pool
        res.append((out.detach().cpu().numpy(), cd.grad.cpu().numpy(), None if pd.grad is None else pd.grad.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0])                                   # fused == unfused, bit for bit
    assert rel_err(res[0][0], ref.detach().numpy()) < 1e-6                        # the host's avg_pool sums in another order
    assert rel_err(res[0][1], res[1][1]) < 1e-6 and rel_err(res[0][1], ct.grad.numpy()) < TOL
    if aa:
        assert np.abs(pt.grad.numpy()).max() > 0
        assert rel_err(res[0][2], pt.grad.numpy()) < TOL and rel_err(res[0][2], res[1][2]) < 1e-5
    else:
        assert res[0][2] is None or float(np.abs(res[0][2]).max()) == 0


@pytest.mark.parametrize("C", [3, 9, 16])
def test_scatter_rows_matches_index_copy(cuda, C):
    """ops.scatter_rows (b2a_rows_scatter / b2a_rows_gather) against torch's new_zeros(...).index_copy / its adjoint: the dense <-> covered
    rows hand-over around the field networks (render._sample_field)."""
    ops = _ops()
    torch.manual_seed(C)
    n, N = 40000, 9000
    idx = torch.randperm(n, device=cuda)[:N].sort().values
    rows = torch.randn(N, C, device=cuda)
    g = torch.randn(n, C, device=cuda)
    a = rows.clone().requires_grad_(True)
    b = rows.clone().requires_grad_(True)
    want = a.new_zeros(n, C).index_copy(0, idx, a)
    got = ops.scatter_rows(b, idx, n)
    assert torch.equal(got, want)
    want.backward(g); got.backward(g)
    assert torch.equal(a.grad, b.grad)
    assert ops.scatter_rows(torch.zeros(0, C, device=cuda), idx[:0], 16).abs().sum().item() == 0


# ----------------------------------------------------------------------------------------------------------------
# fused g-buffer
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("spp,Bq,two_sided,use_cov", [(1, 1, True, True), (1, 1, True, False), (1, 3, False, True), (2, 1, True, False)])
def test_gbuffer_oracle(cuda, spp, Bq, two_sided, use_cov):
    """One fused kernel vs the reference's sequence: 4x interpolate + prepare_shading_normal + camera normal."""
    ops = _ops()
    verts, faces, prior, mvp, w2c, campos, clip = _scene()
    H = W = 64
    rast = R.rasterize(clip, faces, (H * spp, W * spp))
    rng = np.random.RandomState(8)
    prior_b = np.stack([prior * np.float32(1 + 0.01 * i) for i in range(Bq)])
    vt = torch.from_numpy(verts).requires_grad_(True)
    qt = torch.from_numpy(prior_b).requires_grad_(True)
    ct = torch.from_numpy(clip).requires_grad_(True)
    wt = torch.from_numpy(w2c).requires_grad_(True)
    pt = torch.from_numpy(campos).requires_grad_(True)
    ft = torch.from_numpy(faces).long()
    nt = T.auto_normals(vt.detach(), ft).requires_grad_(True)
    # reference sequence on the CPU oracle
    rast_full = T._Rasterize.apply(ct, torch.from_numpy(faces), (H * spp, W * spp))
    assert np.array_equal(rast_full.detach().numpy(), rast)
    rs = rast_full[:, ::spp, ::spp].contiguous()
    tri = torch.from_numpy(faces)
    gb_pos = T.interpolate(vt, rs, tri)
    v0, v1, v2 = vt[:, ft[:, 0]], vt[:, ft[:, 1]], vt[:, ft[:, 2]]
    fn = T.safe_normalize(torch.cross(v1 - v0, v2 - v0, dim=-1))
    gb_geo = T.interpolate(fn, rs, torch.arange(faces.shape[0])[:, None].repeat(1, 3))
    gb_nrm = T.interpolate(nt, rs, tri)
    gb_tex = T.interpolate(qt, rs, tri)
    gb_shn = T.prepare_shading_normal(gb_pos, pt[:, None, None, :], gb_nrm, None, gb_geo, two_sided)
    gb_cam = T.safe_normalize(torch.matmul(gb_shn.view(3, -1, 3), wt[:, :3, :3].transpose(2, 1))).view(3, H, W, 3)
    refs = dict(pos=gb_pos, geo_nrm=gb_geo, shading_nrm=gb_shn, cam_nrm=gb_cam, tex_pos=gb_tex)
    gs = {k: rng.randn(3, H, W, 3).astype(np.float32) for k in refs}
    sum((refs[k] * torch.from_numpy(gs[k])).sum() for k in refs).backward()
    # fused kernel
    vd, nd, qd = (dev(x.detach().numpy(), cuda).requires_grad_(True) for x in (vt, nt, qt))
    cd, wd, pd = (dev(x.detach().numpy(), cuda).requires_grad_(True) for x in (ct, wt, pt))
    coverage = None
    if use_cov:   # the rasterizer's compact covered-pixel list (dense-warp backward)
        rast_d, coverage = ops.rasterize(cd.detach(), dev(faces, cuda), (H, W), with_coverage=True)
        assert np.array_equal(rast_d.cpu().numpy(), rast)
        n_cov = int(coverage[1].item())
        assert n_cov == int((rast[..., 3] > 0).sum())
        cl = coverage[0][:n_cov].cpu().numpy()
        order = np.argsort(cl[:, 0])
        flat = np.flatnonzero(rast[..., 3].reshape(-1) > 0)
        assert np.array_equal(cl[order, 0], flat)
        assert np.array_equal(cl[order, 1:], faces[rast[..., 3].reshape(-1)[flat].astype(np.int64) - 1])   # vertex ids of the visible triangle
    out = ops.gbuffer(dev(rast, cuda), cd, dev(faces, cuda), vd, nd, qd, wd, pd, spp=spp, two_sided=two_sided, want=tuple(refs),
                      coverage=coverage)
    covered = rast[:, ::spp, ::spp, 3] > 0
    for k in refs:
        assert rel_err(out[k].detach().cpu().numpy(), refs[k].detach().numpy()) < TOL, k
        assert np.all(out[k].detach().cpu().numpy()[~covered] == 0)
    sum((out[k] * dev(gs[k], cuda)).sum() for k in refs).backward()
    for name, a, b in (("v_pos", vd, vt), ("v_nrm", nd, nt), ("prior", qd, qt), ("clip", cd, ct), ("w2c", wd, wt), ("campos", pd, pt)):
        assert rel_err(a.grad.cpu().numpy(), b.grad.numpy()) < 2e-4, name


# ----------------------------------------------------------------------------------------------------------------
# directional-light shading
# ----------------------------------------------------------------------------------------------------------------
def test_directional_light_golden(cuda):
    """Drop-in DirectionalLight (light MLP in torch, shading in csrc/shade.cu) vs the reference's own class: outputs and the
    gradients to the light MLP input, the 9-channel texture (kd is read in place as its leading 3 channels) and the normal."""
    light_mod = pkg("render.light")
    g = golden("light_directional.npz")
    lgt = light_mod.DirectionalLight(16, 3, 32, intensity_min_max=torch.zeros(2, 2)).to(cuda)
    lgt.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}, strict=True)
    feat, tex, nrm = (dev(g[k], cuda).requires_grad_(True) for k in ("feat", "tex", "nrm"))
    shaded, shading = lgt.shade(feat, tex[..., :3], nrm)
    assert tuple(shading.shape) == tuple(g["shading"].shape)
    assert rel_err(shaded.detach().cpu().numpy(), g["shaded"]) < TOL and rel_err(shading.detach().cpu().numpy(), g["shading"]) < TOL
    ((shaded * dev(g["g_shaded"], cuda)).sum() + (shading * dev(g["g_shading"], cuda)).sum()).backward()
    for a, k in ((feat, "d_feat"), (tex, "d_tex"), (nrm, "d_nrm")):
        assert rel_err(a.grad.cpu().numpy(), g[k]) < TOL, k


@pytest.mark.parametrize("B,H,W,Bl,dense_kd", [(2, 5, 5, 2, True), (3, 16, 24, 1, True), (2, 8, 8, 2, False), (1, 7, 9, 1, False)])
def test_shade_directional_oracle(cuda, B, H, W, Bl, dense_kd):
    """Vector (H*W % 4 == 0) and scalar paths, shared and per-image light, dense kd and kd as a slice of 9 channels."""
    ops = _ops()
    rng = np.random.RandomState(31)
    tex = rng.rand(B, H, W, 9).astype(np.float32)
    nrm = rng.randn(B, H, W, 3).astype(np.float32)
    nrm[:, :1] = 0
    lp = np.concatenate([rng.randn(Bl, 3), rng.rand(Bl, 2) + 0.2], -1).astype(np.float32)
    g1, g2 = rng.randn(B, H, W, 3).astype(np.float32), rng.randn(B, H, W, 1).astype(np.float32)
    tt, nt, lt = (torch.from_numpy(x).requires_grad_(True) for x in (tex, nrm, lp))
    rs, rh = T.directional_shade(lt.expand(B, 5), tt[..., :3], nt)
    ((rs * torch.from_numpy(g1)).sum() + (rh * torch.from_numpy(g2)).sum()).backward()
    td, nd, ld = (dev(x, cuda).requires_grad_(True) for x in (tex, nrm, lp))
    kd = td[..., :3].contiguous() if dense_kd else td[..., :3]
    s, h = ops.shade_directional(kd, nd, ld)
    assert rel_err(s.detach().cpu().numpy(), rs.detach().numpy()) < TOL and rel_err(h.detach().cpu().numpy(), rh.detach().numpy()) < TOL
    ((s * dev(g1, cuda)).sum() + (h * dev(g2, cuda)).sum()).backward()
    for a, b, k in ((td, tt, "tex"), (nd, nt, "nrm"), (ld, lt, "light")):
        assert rel_err(a.grad.cpu().numpy(), b.grad.numpy()) < TOL, k


@pytest.mark.parametrize("squash,C", [(True, 3), (False, 16), (False, 5)])
def test_analytic_field_stand_in(cuda, squash, C):
    """The bench's stand-in pixel shader (csrc/analytic_field.cu) computes exactly the oracle twin's function."""
    ops = _ops()
    rng = np.random.RandomState(41)
    x = (rng.randn(2, 9, 7, 3) * 2).astype(np.float32)
    w = (rng.randn(3, C) * 1.5).astype(np.float32)
    xt = torch.from_numpy(x).requires_grad_(True)
    z = torch.matmul(xt, torch.from_numpy(w))
    ref = torch.cat([torch.sigmoid(z)] * 3, -1) if squash else torch.sin(z)
    g = rng.randn(*ref.shape).astype(np.float32)
    (ref * torch.from_numpy(g)).sum().backward()
    xd = dev(x, cuda).requires_grad_(True)
    out = ops.analytic_field(xd, dev(w, cuda), squash)
    assert rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < 1e-5
    (out * dev(g, cuda)).sum().backward()
    assert rel_err(xd.grad.cpu().numpy(), xt.grad.numpy()) < 1e-5
