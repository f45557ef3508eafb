"""The hot path end to end on synthetic MagicPony-horse inputs: what bench.py times and smoke() runs.

    sdf values on the tet grid -> DMTet extraction -> make_mesh (normals)            [prior shape, batch 1]
      -> estimate_bones -> skinning (LBS) -> make_mesh (normals)                     [B posed instances]
      -> render_mesh(['shaded','dino_pred'])                                         [clip, raster, g-buffer, shade, AA]
      -> backward from upstream image gradients to d_sdf and d_articulation.

It calls the package's public drop-in API (geometry.dmtet / geometry.skinning / render.mesh / render.render), i.e. the
same entry points the reference's predictors and AnimalModel.render use (InstancePredictorBase.py:513-598,
AnimalModel.py:217-258).  The field networks are PyTorch modules: an analytic colour field (M1a: kernels only) or
CoordMLPs of the reference's sizes (M1b; magicpony.yaml:43-51,65-74).
"""
import numpy as np
import torch

from . import synthetic
from .geometry import dmtet as dmtet_mod
from .geometry import skinning as skinning_mod
from .render import mesh as mesh_mod
from .render import render as render_mod


class SyntheticScene:
    """Seeded numpy inputs shared by the CUDA path and the CPU oracle (same bytes on both sides)."""

    def __init__(self, grid_res=128, batch=16, image_res=256, n_body_bones=8, n_leg_bones=3, dino_dim=16, spatial_scale=7.0,
                 sdf_noise=0.01, seed=0, body_bones_mode="z_minmax_y+", bone_y_threshold=None, chain_every_step=False, static_root_bones=False,
                 second_render=False, class_dim=0):
        self.grid_res, self.batch, self.image_res = grid_res, batch, image_res
        self.n_body_bones, self.n_leg_bones, self.dino_dim = n_body_bones, n_leg_bones, dino_dim
        self.num_bones = n_body_bones + 4 * n_leg_bones
        # config switches of the other shipped workloads (BASELINE configs[2], [3]):
        #   body_bones_mode / n_leg_bones = 0 / static_root_bones: train_magicpony_bird.yaml:29-36
        #   bone_y_threshold + kinematic chain every iteration: 3D-Fauna (InstancePredictorFauna.py:20,82-93)
        #   second_render: Fauna's texture-less / light-less ['shaded'] render from a random view (Fauna.py:145-163,454)
        #   class_dim: the class vector fed to the DINO field (Fauna.py:388; CoordMLP extra_feat_dim)
        self.body_bones_mode, self.bone_y_threshold, self.chain_every_step = body_bones_mode, bone_y_threshold, chain_every_step
        self.static_root_bones, self.second_render, self.class_dim = static_root_bones, second_render, class_dim
        v, t = synthetic.kuhn_tet_grid(grid_res)
        self.grid_verts = (v * np.float32(spatial_scale)).astype(np.float32)
        self.tets = t
        self.sdf = synthetic.sdf_horse(self.grid_verts, sigma=sdf_noise, seed=seed)
        rng = np.random.RandomState(seed + 1)
        self.angles = rng.uniform(-0.3, 0.3, size=(batch, 1, self.num_bones, 3)).astype(np.float32)
        if static_root_bones:       # apply_articulation_constraints (InstancePredictorBase.py:437-441): the two root bones do not move
            self.angles[:, :, [n_body_bones // 2 - 1, n_body_bones - 1]] = 0
        self.mvp, self.w2c, self.campos = synthetic.cameras(batch, seed=seed + 3)
        self.mvp2, self.w2c2, self.campos2 = synthetic.cameras(batch, seed=seed + 17) if second_render else (None, None, None)
        self.class_vector = np.random.RandomState(seed + 5).randn(batch, class_dim).astype(np.float32) if class_dim else None
        rng = np.random.RandomState(seed + 2)
        self.feat = rng.randn(batch, 256).astype(np.float32)
        self.light = np.array([0.3, 0.5, 0.8, 0.4, 0.6], np.float32)   # dir(3, normalised below), ambient, diffuse
        self.light[:3] /= np.linalg.norm(self.light[:3])
        self.w_kd = (rng.randn(3, 3) * 1.5).astype(np.float32)
        self.w_dino = (rng.randn(3, dino_dim) * 1.5).astype(np.float32)

    def upstream_grads(self, seed=7):
        rng = np.random.RandomState(seed)
        r = self.image_res
        return ((rng.randn(self.batch, 4, r, r) * 1e-3).astype(np.float32),
                (rng.randn(self.batch, self.dino_dim, r, r) * 1e-3).astype(np.float32))

    def upstream_grad_mask(self, seed=9):
        """Upstream gradient of the second render's mask (Fauna's mask discriminator loss), [B,1,256,256]."""
        return (np.random.RandomState(seed).randn(self.batch, 1, 256, 256) * 1e-3).astype(np.float32)


class AnalyticField(torch.nn.Module):
    """Fixed smooth colour / feature field standing in for the texture and DINO CoordMLPs (M1a):
    texture = cat([sigmoid(x W)] * 3) (9 channels), dino = sin(x W16) - the same function as
    oracle.pipeline_ref.analytic_shader - as one kernel per direction (csrc/analytic_field.cu): M1a measures the hot-path
    kernels, so the stand-in must cost as little as possible (a K=3 torch.matmul is a ~90 us SIMT sgemm plus ~25 us of
    cuBLAS host time, and its autograd graph adds a dozen kernels per step)."""

    def __init__(self, weight, squash):
        super().__init__()
        self.register_buffer("weight", weight)
        self.squash = squash
        self.bsdf = None
        self.dense_only = True   # three flops per pixel: gathering the covered rows would cost more than evaluating everywhere

    def sample(self, x, feat=None):
        from . import ops
        return ops.analytic_field(x, self.weight, self.squash)


class FixedLight(torch.nn.Module):
    """DirectionalLight.shade (reference model/render/light.py:186-193) with fixed light parameters: the same fused
    kernel the drop-in DirectionalLight uses (render/light.py -> ops.shade_directional)."""

    def __init__(self, params):
        super().__init__()
        self.register_buffer("params", params.reshape(1, 5).contiguous())

    def shade(self, feat, kd, normal):
        from . import ops
        return ops.shade_directional(kd, normal, self.params)


class HotPath(torch.nn.Module):
    def __init__(self, scene, device="cuda", mlps=False):
        super().__init__()
        s = self.scene = scene
        dev = torch.device(device)
        self.grid_verts = torch.from_numpy(s.grid_verts).to(dev)
        self.tets = torch.from_numpy(s.tets).to(dev)
        self.dmtet = dmtet_mod.DMTet()
        self.grid = self.dmtet.grid_for(self.tets, self.grid_verts.shape[0])
        self.sdf = torch.nn.Parameter(torch.from_numpy(s.sdf).to(dev)[:, None])
        self.angles = torch.nn.Parameter(torch.from_numpy(s.angles).to(dev))
        self.mvp = torch.from_numpy(s.mvp).to(dev)
        self.w2c = torch.from_numpy(s.w2c).to(dev)
        self.campos = torch.from_numpy(s.campos).to(dev)
        self.feat = torch.from_numpy(s.feat).to(dev)
        self.light = FixedLight(torch.from_numpy(s.light).to(dev))
        self.class_vector = torch.from_numpy(s.class_vector).to(dev) if s.class_vector is not None else None
        if s.second_render:
            self.mvp2, self.w2c2, self.campos2 = (torch.from_numpy(x).to(dev) for x in (s.mvp2, s.w2c2, s.campos2))
        if mlps:
            from .networks import CoordMLP
            mm = torch.tensor([[0., 1.]] * 9, device=dev)
            self.material = CoordMLP(3, 9, 8, nf=256, activation="sigmoid", min_max=mm, n_harmonic_functions=10,
                                     embedder_scalar=2 * np.pi / 7.0 * 0.9, extra_feat_dim=256, symmetrize=True).to(dev)
            self.dino_net = CoordMLP(3, s.dino_dim, 5, nf=256, activation="sigmoid", n_harmonic_functions=8,
                                     embedder_scalar=2 * np.pi / 7.0 * 0.9, extra_feat_dim=s.class_dim, symmetrize=True).to(dev)
        else:
            self.material = AnalyticField(torch.from_numpy(s.w_kd).to(dev), True)
            self.dino_net = AnalyticField(torch.from_numpy(s.w_dino).to(dev), False)
        self.mlps = mlps
        self.sparse_fields = True
        self.spp = 1
        self.bone_aux = None
        self.kinematic_chain = None

    def forward(self, sdf=None, render_modes=("shaded", "dino_pred")):
        s = self.scene
        sdf = self.sdf if sdf is None else sdf
        # prior shape: extraction + normals (DMTetGeometry.getMesh, dmtet.py:294-310)
        verts, faces, uv_idx, faces32 = self.dmtet.extract(self.grid_verts, sdf, self.grid)
        prior = mesh_mod.make_mesh(verts[None], faces[None], None, uv_idx[None], None, faces_i32=faces32)
        # bones from the prior shape (InstancePredictorBase.py:319-335); chain lists only the first time
        ekw = dict(n_legs=4, n_leg_bones=s.n_leg_bones, body_bones_mode=s.body_bones_mode)
        if s.bone_y_threshold is not None:
            ekw["bone_y_threshold"] = s.bone_y_threshold
        if self.bone_aux is None or s.chain_every_step:     # 3D-Fauna recomputes the chain every iteration (InstancePredictorFauna.py:82-93)
            bones, self.kinematic_chain, self.bone_aux = skinning_mod.estimate_bones(
                prior.v_pos[:, None].detach(), s.n_body_bones, compute_kinematic_chain=True, **ekw)
        else:
            bones = skinning_mod.estimate_bones(prior.v_pos[:, None].detach(), s.n_body_bones, compute_kinematic_chain=False,
                                                aux=self.bone_aux, **ekw)
        # articulation (InstancePredictorBase.py:578-586)
        posed, aux = skinning_mod.skinning(prior.v_pos[:, None], bones, self.kinematic_chain, self.angles, output_posed_bones=True,
                                           temperature=0.05)
        inst = mesh_mod.make_mesh(posed[:, 0], prior.t_pos_idx, None, prior.t_tex_idx, None, faces_i32=prior.tri_i32())
        inst._opp = prior._opp
        res = (s.image_res, s.image_res)
        feat = self.feat if self.mlps else None
        out = render_mod.render_mesh(None, inst, self.mvp, self.w2c, self.campos, self.material, self.light, res, spp=self.spp,
                                     num_layers=1, msaa=True, background=None, bsdf="diffuse", feat=feat,
                                     render_modes=list(render_modes), prior_mesh=prior, dino_net=self.dino_net,
                                     class_vector=self.class_vector if self.mlps else None, sparse_fields=self.sparse_fields)
        if s.second_render:     # FaunaModel.get_random_view_mask (Fauna.py:145-163): no texture, no light, one-sided, 256^2
            out = list(out) + [render_mod.render_mesh(None, inst, self.mvp2, self.w2c2, self.campos2, None, None, (256, 256), spp=1, num_layers=1,
                                                      msaa=True, background=None, bsdf="diffuse", feat=None, render_modes=["shaded"],
                                                      prior_mesh=prior, two_sided_shading=False, dino_net=None)[0][:, 3:]]
        self.last = dict(prior=prior, inst=inst, bones=bones, posed_bones=aux["posed_bones"])
        return out

    def step(self, d_shaded, d_dino, d_mask=None):
        """One fwd+bwd pass; returns (d_sdf [Vg,1], d_angles [B,1,K,3]).  d_mask: upstream gradient of the second render's mask."""
        self.sdf.grad = None
        self.angles.grad = None
        outs = self.forward()
        torch.autograd.backward(list(outs), [d_shaded, d_dino] + ([d_mask] if len(outs) > 2 else []))
        return self.sdf.grad, self.angles.grad
