"""CPU twin of 3danimals_b200.pipeline.HotPath (TEST INFRASTRUCTURE ONLY).

Runs the same hot path on host cores from the same SyntheticScene bytes: the restated reference geometry
(oracle.torch_ref / oracle.geometry_np, pinned against the reference's own files) plus the C restatement of the
nvdiffrast ops (oracle/raster_ref.c).  Used by tests/ and smoke() as the parity checker and by bench.py as the
`cpu_baseline` / `--impl reference` arm.  Never imported by the product.
"""
import torch

from . import geometry_np as gnp
from . import torch_ref as T


def analytic_shader(scene, images):
    w_kd = torch.from_numpy(scene.w_kd)
    w_dino = torch.from_numpy(scene.w_dino)
    light = torch.from_numpy(scene.light)

    def shade(gb_tex, cam_normal, gbuf):
        kd = torch.sigmoid(torch.matmul(gb_tex, w_kd))
        shading = light[3] + light[4] * torch.clamp(torch.sum(light[:3] * cam_normal, -1, keepdim=True), min=0.0)
        return {"shaded": shading * kd, "kd": kd, "shading": shading, "dino_pred": torch.sin(torch.matmul(gb_tex, w_dino))}

    return shade


def forward(scene, images=None, sdf=None, angles=None, render_modes=("shaded", "dino_pred"), spp=1, background=None,
            fast_normals=True):
    """-> dict(outputs..., sdf, angles leaf tensors, verts, faces, bones, kinematic_chain, posed, rast, ...)."""
    B = scene.batch if images is None else images
    pos = torch.from_numpy(scene.grid_verts)
    tets = torch.from_numpy(scene.tets)
    sdf = torch.from_numpy(scene.sdf)[:, None].clone().requires_grad_(True) if sdf is None else sdf
    angles = torch.from_numpy(scene.angles[:B]).clone().requires_grad_(True) if angles is None else angles
    normals = T.auto_normals_c if fast_normals else T.auto_normals
    verts, faces, uv_idx = T.marching_tets(pos, sdf, tets)
    bones, chain, aux = gnp.estimate_bones(verts.detach().numpy()[None, None], scene.n_body_bones, n_legs=4,
                                           n_leg_bones=scene.n_leg_bones, body_bones_mode=getattr(scene, "body_bones_mode", "z_minmax_y+"),
                                           bone_y_threshold=getattr(scene, "bone_y_threshold", None))
    bones_t = torch.from_numpy(bones)
    posed, saux = T.skinning(verts[None, None], bones_t, chain, angles, temperature=0.05)
    v_nrm = normals(posed[:, 0], faces)
    r = scene.image_res
    out = T.render_mesh(posed[:, 0], v_nrm, faces, torch.from_numpy(scene.mvp[:B]), torch.from_numpy(scene.w2c[:B]),
                        torch.from_numpy(scene.campos[:B]), analytic_shader(scene, B), (r, r), spp=spp, background=background,
                        render_modes=render_modes, prior_v_pos=verts[None])
    if getattr(scene, "second_render", False):      # FaunaModel.get_random_view_mask (Fauna.py:145-163)
        def plain(gb_tex, cam_normal, gbuf):
            kd = torch.ones(*gb_tex.shape[:-1], 3)
            return {"shaded": kd, "kd": kd}
        o2 = T.render_mesh(posed[:, 0], v_nrm, faces, torch.from_numpy(scene.mvp2[:B]), torch.from_numpy(scene.w2c2[:B]),
                           torch.from_numpy(scene.campos2[:B]), plain, (256, 256), spp=1, background=None, render_modes=("shaded",),
                           prior_v_pos=verts[None], two_sided_shading=False)
        out["mask2"] = o2["shaded"][:, 3:]
    out.update(sdf=sdf, angles=angles, verts=verts, faces=faces, uv_idx=uv_idx, bones=bones_t, kinematic_chain=chain,
               posed=posed, posed_bones=saux["posed_bones"], v_nrm=v_nrm)
    return out


def step(scene, d_shaded, d_dino, images=None, d_mask=None):
    """One fwd+bwd pass on the CPU; returns (d_sdf, d_angles, outputs)."""
    B = scene.batch if images is None else images
    out = forward(scene, images=B)
    outs, grads = [out["shaded"], out["dino_pred"]], [torch.from_numpy(d_shaded[:B]), torch.from_numpy(d_dino[:B])]
    if "mask2" in out and d_mask is not None:
        outs.append(out["mask2"]); grads.append(torch.from_numpy(d_mask[:B]))
    torch.autograd.backward(outs, grads)
    return out["sdf"].grad, out["angles"].grad, out
