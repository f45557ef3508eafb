"""The CPU oracle against the committed golden vectors (generated from the reference's own Python by
tests/golden/make_goldens.py).  This is what pins the oracle everywhere, including the GPU box."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_files, pkg, rel_err
from oracle import geometry_np as gnp
from oracle import torch_ref as T

syn = pkg("synthetic")


@pytest.mark.parametrize("name", golden_files("mt_"))
def test_marching_tets_matches_reference(name):
    g = golden(name)
    v, t = syn.kuhn_tet_grid(int(g["res"]))
    v = v * np.float32(7.0)
    o = gnp.marching_tets(v, g["sdf"], t, with_uvs=True)
    assert np.array_equal(o["faces"], g["faces"])          # bit-exact index buffers
    assert np.array_equal(o["uv_idx"], g["uv_idx"])
    assert o["verts"].shape == g["verts"].shape
    assert rel_err(o["verts"], g["verts"]) < 1e-6
    assert tuple(o["uvs"].shape) == tuple(g["uvs_shape"])
    assert np.allclose(o["uvs"][:64], g["uvs_head"], atol=1e-7)
    d_sdf, _ = gnp.lerp_vertices_bwd(v, g["sdf"], o["interp_v"], g["d_verts"])
    assert rel_err(d_sdf, g["d_sdf"]) < 1e-4


def _chain(g):
    return [(int(b), [int(x) for x in str(d).split(",") if x != ""]) for b, d in zip(g["chain_ids"], g["chain_dep"])]


@pytest.mark.parametrize("name", golden_files("skin_"))
def test_bones_and_skinning_match_reference(name):
    g = golden(name)
    verts = g["verts"]
    n_leg, mode = int(g["n_leg_bones"]), str(g["mode"])
    bones, chain, aux = gnp.estimate_bones(verts[None, None], 8, n_legs=4, n_leg_bones=n_leg, body_bones_mode=mode)
    assert chain == _chain(g)
    assert np.allclose(bones, g["bones"], atol=1e-6)
    bones2 = gnp.estimate_bones(verts[None, None] * np.float32(1.01), 8, n_legs=4, n_leg_bones=n_leg, body_bones_mode=mode,
                                compute_kinematic_chain=False, aux=aux)
    assert np.allclose(bones2, g["bones_rescaled"], atol=1e-6)
    out, w, posed = gnp.skinning(verts[None, None], g["bones"], chain, g["angles"], temperature=0.05)
    assert rel_err(out, g["out"]) < 1e-5
    assert rel_err(posed, g["posed_bones"]) < 1e-5
    assert np.abs(w - g["weights"]).max() < 1e-5
    # torch twin incl. gradients
    ang = torch.from_numpy(g["angles"]).requires_grad_(True)
    vp = torch.from_numpy(verts)[None, None].clone().requires_grad_(True)
    o2, aux2 = T.skinning(vp, torch.from_numpy(g["bones"]), chain, ang, temperature=0.05)
    ((o2 * torch.from_numpy(g["g_out"])).sum() + (aux2["posed_bones"] * torch.from_numpy(g["g_posed"])).sum()).backward()
    assert rel_err(o2.detach().numpy(), g["out"]) < 1e-5
    assert rel_err(ang.grad.numpy(), g["d_angles"]) < 1e-4
    assert rel_err(vp.grad.numpy(), g["d_verts"]) < 1e-4


def test_shading_normal_matches_reference():
    g = golden("shading_normal.npz")
    t = lambda k: torch.from_numpy(g[k]).requires_grad_(True)
    pos, nrm, geo = t("pos"), t("nrm"), t("geo")
    for tng in (torch.from_numpy(g["tng"]), None):  # the tangent is numerically dead
        out = T.prepare_shading_normal(pos, torch.from_numpy(g["view"]), nrm, tng, geo, True)
        assert rel_err(out.detach().numpy(), g["out"]) < 1e-6
    (out * torch.from_numpy(g["g"])).sum().backward()
    assert rel_err(pos.grad.numpy(), g["d_pos"]) < 1e-4
    assert rel_err(nrm.grad.numpy(), g["d_nrm"]) < 1e-4
    assert rel_err(geo.grad.numpy(), g["d_geo"]) < 1e-4


@pytest.mark.parametrize("case", ["mesh", "degenerate"])
def test_vertex_normals_match_reference(case):
    """R3: every restatement of `auto_normals` (torch ops, the C twin used by the CPU baseline, numpy) against the output and the
    gradient of the reference's own function (model/render/mesh.py:276-304; tests/golden/normals.npz), incl. the (0,0,1)
    fallback of a vertex no face uses and of one touched by zero-area faces only."""
    g = golden("normals.npz")
    faces = torch.from_numpy(g[case + "_faces"]).long()
    for fn in (T.auto_normals, T.auto_normals_c):
        v = torch.from_numpy(g[case + "_v_pos"]).requires_grad_(True)
        nrm = fn(v, faces)
        assert rel_err(nrm.detach().numpy(), g[case + "_v_nrm"]) < 1e-6, fn.__name__
        (nrm * torch.from_numpy(g[case + "_g"])).sum().backward()
        assert rel_err(v.grad.numpy(), g[case + "_d_v_pos"]) < 1e-5, fn.__name__
    nn = gnp.auto_normals(g[case + "_v_pos"], g[case + "_faces"])
    assert rel_err(nn, g[case + "_v_nrm"]) < 1e-6


def test_directional_shade_matches_reference():
    """oracle.torch_ref.directional_shade vs the reference's DirectionalLight.shade (light.py:186-193) incl. gradients."""
    g = golden("light_directional.npz")
    tex = torch.from_numpy(g["tex"]).requires_grad_(True)
    nrm = torch.from_numpy(g["nrm"]).requires_grad_(True)
    lp = torch.from_numpy(g["light_params"])
    shaded, shading = T.directional_shade(lp, tex[..., :3], nrm)
    assert rel_err(shaded.detach().numpy(), g["shaded"]) < 1e-6 and rel_err(shading.detach().numpy(), g["shading"]) < 1e-6
    ((shaded * torch.from_numpy(g["g_shaded"])).sum() + (shading * torch.from_numpy(g["g_shading"])).sum()).backward()
    assert rel_err(tex.grad.numpy(), g["d_tex"]) < 1e-5
    assert rel_err(nrm.grad.numpy(), g["d_nrm"]) < 1e-5


@pytest.mark.parametrize("name", golden_files("mt_")[:4] + golden_files("skin_"))
def test_torch_op_geometry_baseline_matches_reference(name):
    """oracle/torch_ops_geometry.py (the reference's torch-op formulation, timed on the GPU by bench.py as the second
    baseline of SURVEY §8d) reproduces the reference-generated goldens on CPU."""
    from oracle import torch_ops_geometry as G
    g = golden(name)
    if name.startswith("mt_"):
        v, t = syn.kuhn_tet_grid(int(g["res"]))
        pos = torch.from_numpy(v * np.float32(7.0))
        sdf = torch.from_numpy(g["sdf"]).requires_grad_(True)
        verts, faces = G.marching_tets(pos, sdf, torch.from_numpy(t).long())
        assert np.array_equal(faces.numpy(), g["faces"])
        assert rel_err(verts.detach().numpy(), g["verts"]) < 1e-6
        (verts * torch.from_numpy(g["d_verts"])).sum().backward()
        assert rel_err(sdf.grad.numpy().reshape(-1), g["d_sdf"].reshape(-1)) < 1e-4
        n = G.auto_normals(verts.detach()[None], faces)
        assert rel_err(n.numpy(), gnp.auto_normals(verts.detach().numpy()[None], faces.numpy())) < 1e-5
    else:
        ang = torch.from_numpy(g["angles"]).requires_grad_(True)
        vp = torch.from_numpy(g["verts"])[None, None].clone().requires_grad_(True)
        out = G.skinning(vp, torch.from_numpy(g["bones"]), _chain(g), ang, temperature=0.05)
        assert rel_err(out.detach().numpy(), g["out"]) < 1e-5
        (out * torch.from_numpy(g["g_out"])).sum().backward()
        # the golden's d_angles also carries the posed-bones term; compare against the torch twin with that term removed
        ang2 = torch.from_numpy(g["angles"]).requires_grad_(True)
        vp2 = torch.from_numpy(g["verts"])[None, None].clone().requires_grad_(True)
        o2, _ = T.skinning(vp2, torch.from_numpy(g["bones"]), _chain(g), ang2, temperature=0.05)
        (o2 * torch.from_numpy(g["g_out"])).sum().backward()
        assert rel_err(ang.grad.numpy(), ang2.grad.numpy()) < 1e-4
        assert rel_err(vp.grad.numpy(), vp2.grad.numpy()) < 1e-4


@pytest.mark.parametrize("tag", ["single", "batch"])
def test_fauna_bones_variant_matches_reference(tag):
    """bone_y_threshold = 0.4 (3D-Fauna, InstancePredictorFauna.py:20,84-99; skinning.py:163-175): numpy oracle vs the reference."""
    g = golden("bones_fauna.npz")
    shape = g[tag + "_shape"]
    bones, chain, aux = gnp.estimate_bones(shape, 8, n_legs=4, n_leg_bones=3, body_bones_mode="z_minmax_y+", bone_y_threshold=0.4)
    ref_chain = [(int(b), [int(x) for x in str(d).split(",") if x != ""]) for b, d in zip(g[tag + "_chain_ids"], g[tag + "_chain_dep"])]
    assert chain == ref_chain
    assert [l["body_bone_idx"] for l in aux["legs"]] == list(g[tag + "_attach"])
    assert np.allclose(bones, g[tag + "_bones"], atol=1e-6)
    bones2 = gnp.estimate_bones(shape * np.float32(1.01), 8, n_legs=4, n_leg_bones=3, body_bones_mode="z_minmax_y+",
                                compute_kinematic_chain=False, aux=aux, bone_y_threshold=0.4)
    assert np.allclose(bones2, g[tag + "_bones_rescaled"], atol=1e-6)
