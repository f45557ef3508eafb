"""torch-CPU restatement of the hot path with autograd (TEST INFRASTRUCTURE ONLY).

This is the "what the reference does" pipeline executed on host cores: the reference's own torch-op geometry
(restated; pinned against the real files by tests/test_oracle_vs_reference.py) plus the C restatement of the
nvdiffrast ops (oracle/raster_ref.c) wrapped as autograd Functions.  It is the parity checker for the CUDA
product and the CPU baseline that bench.py times; it is never imported by the product.
Each function cites the reference file:line it follows (relative to /root/reference).
"""
import torch

from . import geometry_np as gnp
from . import raster as R


# ---------------------------------------------------------------------------------------------
# helpers (render/util.py:22-32)
# ---------------------------------------------------------------------------------------------
def dot(x, y):
    return torch.sum(x * y, -1, keepdim=True)


def safe_normalize(x, eps=1e-20):
    return x / torch.sqrt(torch.clamp(dot(x, x), min=eps))


# ---------------------------------------------------------------------------------------------
# R2 marching tets: index work in numpy, differentiable lerp in torch (dmtet.py:104-155)
# ---------------------------------------------------------------------------------------------
def marching_tets(pos, sdf, tets):
    o = gnp.marching_tets(pos.detach().numpy(), sdf.detach().numpy().reshape(-1), tets.numpy(), with_uvs=False)
    iv = torch.from_numpy(o["interp_v"])
    s = sdf.reshape(-1)
    sa, sb = s[iv[:, 0]], -s[iv[:, 1]]
    den = sa + sb
    verts = pos[iv[:, 0]] * (sb / den)[:, None] + pos[iv[:, 1]] * (sa / den)[:, None]
    return verts, torch.from_numpy(o["faces"]), torch.from_numpy(o["uv_idx"])


# ---------------------------------------------------------------------------------------------
# R3 vertex normals (render/mesh.py:276-304), torch ops exactly as the reference minus the 'cuda' literals
# ---------------------------------------------------------------------------------------------
def auto_normals(v_pos, faces):
    B = v_pos.shape[0]
    i0, i1, i2 = faces[:, 0], faces[:, 1], faces[:, 2]
    v0, v1, v2 = v_pos[:, i0], v_pos[:, i1], v_pos[:, i2]
    fn = torch.cross(v1 - v0, v2 - v0, dim=-1)
    v_nrm = torch.zeros_like(v_pos)
    for idx in (i0, i1, i2):
        v_nrm = v_nrm.scatter_add(1, idx[None, :, None].repeat(B, 1, 3), fn)
    v_nrm = torch.where(dot(v_nrm, v_nrm) > 1e-20, v_nrm, torch.tensor([0.0, 0.0, 1.0]))
    return safe_normalize(v_nrm)


class _NormalsC(torch.autograd.Function):
    """Same arithmetic through the C restatement (fast; used by the CPU baseline)."""

    @staticmethod
    def forward(ctx, v_pos, faces):
        nrm, nsum = R.vertex_normals(v_pos.detach().numpy(), faces.numpy())
        ctx.save_for_backward(v_pos, faces, torch.from_numpy(nsum))
        return torch.from_numpy(nrm)

    @staticmethod
    def backward(ctx, g):
        v_pos, faces, nsum = ctx.saved_tensors
        return torch.from_numpy(R.vertex_normals_bwd(v_pos.detach().numpy(), faces.numpy(), nsum.numpy(),
                                                     g.contiguous().numpy())), None


def auto_normals_c(v_pos, faces):
    return _NormalsC.apply(v_pos, faces)


# ---------------------------------------------------------------------------------------------
# R5 skinning (skinning.py:369-439) in closed form (verified against the reference)
# ---------------------------------------------------------------------------------------------
def _euler_xyz(a):
    x, y, z = a.unbind(-1)
    cx, sx, cy, sy, cz, sz = x.cos(), x.sin(), y.cos(), y.sin(), z.cos(), z.sin()
    o, n = torch.ones_like(x), torch.zeros_like(x)
    Rx = torch.stack([o, n, n, n, cx, -sx, n, sx, cx], -1).reshape(x.shape + (3, 3))
    Ry = torch.stack([cy, n, sy, n, o, n, -sy, n, cy], -1).reshape(x.shape + (3, 3))
    Rz = torch.stack([cz, -sz, n, sz, cz, n, n, n, o], -1).reshape(x.shape + (3, 3))
    return Rx @ Ry @ Rz


def _rest_frames(bones):
    joint = bones[..., 0, :]
    fwd = torch.nn.functional.normalize(bones[..., 1, :] - bones[..., 0, :], p=2, dim=-1)     # :257
    right = torch.tensor([1.0, 0.0, 0.0]).expand_as(fwd)
    up = torch.nn.functional.normalize(torch.cross(fwd, right, dim=-1), p=2, dim=-1)           # :261-262
    right = torch.cross(up, fwd, dim=-1)                                                      # :263
    Rm = torch.stack([right, up, fwd], -1)                                                    # :266
    M = torch.zeros(bones.shape[:-2] + (4, 4))
    M[..., :3, :3] = Rm
    M[..., :3, 3] = joint
    M[..., 3, 3] = 1
    Mi = torch.zeros_like(M)
    Mi[..., :3, :3] = Rm.transpose(-1, -2)
    Mi[..., :3, 3] = -(Rm.transpose(-1, -2) @ joint[..., None])[..., 0]
    Mi[..., 3, 3] = 1
    return M, Mi


def bone_transforms(bones, angles, kinematic_tree):
    B, F, K = angles.shape[:3]
    bones = bones.expand(B, F, *bones.shape[2:])
    Rest, RestInv = _rest_frames(bones)
    Rot = torch.zeros(B, F, K, 4, 4)
    Rot[..., :3, :3] = _euler_xyz(angles)
    Rot[..., 3, 3] = 1
    T = Rest @ Rot @ RestInv
    G = [None] * K
    for k, chain in gnp.chain_lists(kinematic_tree).items():
        M = T[:, :, chain[0]]
        for i in chain[1:]:
            M = M @ T[:, :, i]
        G[k] = M
    return torch.stack(G, 2)


def skinning_weights(bones, v_pos, temperature):
    a, b = bones[:, :, :, 0, None, :], bones[:, :, :, 1, None, :]           # [Bb,Fb,K,1,3]
    p = v_pos[:, :, None]                                                    # [Bv,Fv,1,V,3]
    ab = b - a
    t = ((p - a) * ab).sum(-1, keepdim=True) / torch.clamp((ab * ab).sum(-1, keepdim=True), min=1e-6)
    s = a + t.clamp(0.0, 1.0) * ab
    d = torch.sqrt(((s - p) ** 2).sum(-1) + 1e-6)                            # [B,F,K,V]
    return torch.softmax(-d / temperature, dim=2)


def skinning(v_pos, bones, kinematic_tree, angles, temperature=1.0):
    """Returns verts [B,F,V,3], aux{vertices_to_bones [K,B',F',V], posed_bones [B,F,K,2,3]}."""
    B, F, K = angles.shape[:3]
    w = skinning_weights(bones, v_pos.detach(), temperature)                # detached verts (skinning.py:377)
    G = bone_transforms(bones, angles, kinematic_tree)                      # [B,F,K,4,4]
    v4 = torch.cat([v_pos, torch.ones_like(v_pos[..., :1])], -1).expand(B, F, -1, -1)
    xk = torch.einsum("bfkij,bfvj->bfkvi", G, v4)[..., :3]
    out = (w[..., None] * xk).sum(2)
    b4 = torch.cat([bones.expand(B, F, -1, -1, -1), torch.ones(B, F, K, 2, 1)], -1)
    posed = torch.einsum("bfkij,bfkej->bfkei", G, b4)[..., :3]
    return out, dict(vertices_to_bones=w.permute(2, 0, 1, 3), posed_bones=posed)


# ---------------------------------------------------------------------------------------------
# R6/R7/R9: nvdiffrast-op restatements (C) as autograd Functions
# ---------------------------------------------------------------------------------------------
def xfm_points(points, matrix):
    """renderutils/ops.py:524-525 (use_python=True branch)."""
    return torch.matmul(torch.nn.functional.pad(points, pad=(0, 1), mode="constant", value=1.0),
                        torch.transpose(matrix, 1, 2))


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, tri, resolution):
        rast = torch.from_numpy(R.rasterize(pos.detach().numpy(), tri.numpy(), resolution))
        ctx.save_for_backward(pos, tri, rast)
        return rast

    @staticmethod
    def backward(ctx, g):
        pos, tri, rast = ctx.saved_tensors
        return torch.from_numpy(R.rasterize_bwd(pos.detach().numpy(), tri.numpy(), rast.numpy(),
                                                g.contiguous().numpy())), None, None


class _Interpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tri):
        ctx.save_for_backward(attr, rast, tri)
        return torch.from_numpy(R.interpolate(attr.detach().numpy(), rast.detach().numpy(), tri.numpy()))

    @staticmethod
    def backward(ctx, g):
        attr, rast, tri = ctx.saved_tensors
        da, dr = R.interpolate_bwd(attr.detach().numpy(), rast.detach().numpy(), tri.numpy(), g.contiguous().numpy())
        return torch.from_numpy(da), torch.from_numpy(dr), None


class _Antialias(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, rast, pos, tri, opp):
        ctx.save_for_backward(color, rast, pos, tri, opp)
        return torch.from_numpy(R.antialias(color.detach().numpy(), rast.detach().numpy(), pos.detach().numpy(),
                                            tri.numpy(), opp.numpy()))

    @staticmethod
    def backward(ctx, g):
        color, rast, pos, tri, opp = ctx.saved_tensors
        dc, dp = R.antialias_bwd(color.detach().numpy(), rast.detach().numpy(), pos.detach().numpy(), tri.numpy(),
                                 g.contiguous().numpy(), opp.numpy())
        return torch.from_numpy(dc), None, torch.from_numpy(dp), None, None


def rasterize(pos, tri, resolution):
    return _Rasterize.apply(pos.contiguous(), tri.int().contiguous(), tuple(resolution))


def interpolate(attr, rast, tri):
    return _Interpolate.apply(attr.contiguous(), rast.contiguous(), tri.int().contiguous())


def edge_adjacency(tri, V):
    return torch.from_numpy(R.edge_adjacency(tri.int().numpy(), V))


def antialias(color, rast, pos, tri, opp=None):
    tri = tri.int().contiguous()
    if opp is None:
        opp = edge_adjacency(tri, pos.shape[1])
    return _Antialias.apply(color.contiguous(), rast.contiguous(), pos.contiguous(), tri, opp)


# ---------------------------------------------------------------------------------------------
# R8 shading normal (renderutils/bsdf.py:25-51 with perturbed_nrm = (0,0,1), ops.py:217-218)
# ---------------------------------------------------------------------------------------------
def prepare_shading_normal(pos, view_pos, smooth_nrm, smooth_tng, geom_nrm, two_sided_shading=True):
    nrm = torch.nn.functional.normalize(smooth_nrm, dim=-1)
    view = torch.nn.functional.normalize(view_pos - pos, dim=-1)
    if smooth_tng is not None:  # reference path: tng*0 - bitang*0 + nrm*1, then normalize (bsdf.py:38-44)
        tng = torch.nn.functional.normalize(smooth_tng, dim=-1)
        bitang = torch.nn.functional.normalize(torch.cross(tng, nrm, dim=-1), dim=-1)
        shading = tng * 0.0 - bitang * 0.0 + nrm * 1.0
    else:
        shading = nrm
    shading = torch.nn.functional.normalize(shading, dim=-1)
    if two_sided_shading:                                                                      # bsdf.py:30-32
        front = dot(geom_nrm, view) > 0
        shading = torch.where(front, shading, -shading)
        geom_nrm = torch.where(front, geom_nrm, -geom_nrm)
    t = torch.clamp(dot(view, shading) / 0.1, min=0, max=1)                                    # bsdf.py:34
    return torch.lerp(geom_nrm, shading, t)


def directional_shade(light_params, kd, cam_normal):
    """light.DirectionalLight.shade given forward()'s light_params [B,5] (light.py:186-193)."""
    ldir = light_params[..., :3][:, None, None, :]
    amb = light_params[..., 3:4][:, None, None, :]
    diff = light_params[..., 4:5][:, None, None, :]
    shading = amb + diff * torch.clamp(dot(ldir, cam_normal), min=0.0)
    return shading * kd, shading


# ---------------------------------------------------------------------------------------------
# render_mesh restated end to end (render/render.py:139-337), spp=1 and spp>1 (msaa) paths
# ---------------------------------------------------------------------------------------------
def _scale_nearest(x, size):
    y = torch.nn.functional.interpolate(x.permute(0, 3, 1, 2), size, mode="nearest")
    return y.permute(0, 2, 3, 1).contiguous()


def render_mesh(v_pos, v_nrm, faces, mtx, w2c, view_pos, shade_fn, resolution, spp=1, background=None,
                render_modes=("shaded",), prior_v_pos=None, two_sided_shading=True, opp=None, num_frames=None):
    """shade_fn(gb_tex_pos, cam_normal, gbuffers) -> dict mode -> [B,h,w,C] colour (without alpha).

    Returns dict mode -> [B,C',H,W] like render.render_mesh, plus 'rast' and 'v_pos_clip'."""
    B = mtx.shape[0]
    H, W = resolution
    full = (H * spp, W * spp)
    tri = faces.int()
    if prior_v_pos is None:
        prior_v_pos = v_pos
    clip = xfm_points(v_pos, mtx)                                                              # render.py:278
    rast = rasterize(clip, tri, full)                                                          # :292-294
    rast_s = rast[:, ::spp, ::spp].contiguous() if spp > 1 else rast                           # :170-172 (nearest)
    gb_pos = interpolate(v_pos, rast_s, tri)                                                   # :182
    v0, v1, v2 = v_pos[:, faces[:, 0]], v_pos[:, faces[:, 1]], v_pos[:, faces[:, 2]]
    face_normals = safe_normalize(torch.cross(v1 - v0, v2 - v0, dim=-1))                       # :185-188
    fidx = torch.arange(faces.shape[0])[:, None].repeat(1, 3)
    gb_geo = interpolate(face_normals, rast_s, fidx)                                           # :189-191
    gb_nrm = interpolate(v_nrm, rast_s, tri)                                                   # :195
    gb_tex = interpolate(prior_v_pos, rast_s, tri)                                             # :209
    vp = view_pos[:, None, None, :] if view_pos.dim() == 2 else view_pos
    gb_shn = prepare_shading_normal(gb_pos, vp, gb_nrm, None, gb_geo, two_sided_shading)       # :72
    b, h, w, _ = gb_shn.shape
    cam_normal = safe_normalize(torch.matmul(gb_shn.view(b, -1, 3), w2c[:, :3, :3].transpose(2, 1))).view(b, h, w, 3)
    gbuf = dict(gb_pos=gb_pos, gb_geo=gb_geo, gb_nrm=gb_nrm, gb_tex=gb_tex, gb_shn=gb_shn, cam_normal=cam_normal)
    if "flow" in render_modes:                                                                 # :281-288
        c2 = clip[..., :2] / clip[..., -1:]
        c2 = c2.view(-1, num_frames, *c2.shape[1:])
        dxy = c2[:, 1:] - c2[:, :-1]
        dxy = torch.cat([dxy, torch.zeros_like(dxy[:, :1])], 1).view(-1, *c2.shape[2:])
        gbuf["flow"] = interpolate(dxy, rast_s, tri)
    buffers = shade_fn(gb_tex, cam_normal, gbuf)
    buffers.setdefault("geo_normal", (gb_geo + 1.0) * 0.5)
    buffers.setdefault("normal", (gb_shn + 1.0) * 0.5)
    if "flow" in render_modes:
        buffers["flow"] = gbuf["flow"]
    if background is not None:                                                                 # :298-304
        bgf = _scale_nearest(background, full) if spp > 1 else background
        bgf = torch.cat((bgf, torch.zeros_like(bgf[..., 0:1])), -1)
    else:
        bgf = torch.zeros(1, full[0], full[1], 4)
    out = {"rast": rast, "v_pos_clip": clip, "gbuf": gbuf}
    if opp is None:
        opp = edge_adjacency(tri, v_pos.shape[1])
    for key in render_modes:
        col = buffers[key]
        if spp > 1:
            col = _scale_nearest(col, full)                                                    # :217-219
        aa = key in ("shaded", "flow", "dino_pred", "depth", "shading")                        # :311
        bg = bgf if key in ("shaded", "geo_normal", "shading") else torch.zeros(*col.shape[:-1], col.shape[-1] + 1)
        if key == "shading":                                                                   # :313-314 (always true there:
            bg = bg[..., 2:]                                                                   #  `background` was reassigned :304)
        alpha = (rast[..., -1:] > 0).float()                                                   # :261
        accum = torch.lerp(bg.expand(B, -1, -1, -1), torch.cat((col, torch.ones_like(col[..., :1])), -1), alpha)
        if aa:
            accum = antialias(accum.contiguous(), rast, clip, tri, opp)                        # :264
        if spp > 1:
            accum = torch.nn.functional.avg_pool2d(accum.permute(0, 3, 1, 2), spp).permute(0, 2, 3, 1)
        if key in ("kd", "ks", "normal", "geo_normal"):
            accum = accum[..., :3]
        elif key in ("shading", "depth"):
            accum = accum[..., :1]
        elif key == "flow":
            accum = accum[..., :2]
        elif key == "dino_pred":
            accum = accum[..., :-1]
        out[key] = accum.permute(0, 3, 1, 2)
    return out
