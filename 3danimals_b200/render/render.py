"""Drop-in for the reference's model/render/render.py: render_mesh (+ shade / render_layer folded in) and render_uv.

Same signature, same list-of-NCHW return (reference render.py:228-337).  What changed underneath:
  * clip transform, rasterizer, g-buffer interpolation + shading normal + camera normal, composite + antialias are
    sm_100a kernels of libb2a.so (ops.xfm_points / rasterize / gbuffer / composite_antialias) - no nvdiffrast;
  * the five dr.interpolate calls and ~20 elementwise shading-normal kernels of render_layer/shade (render.py:182-209,
    :72-75) are ONE fused launch; only gb_tex_pos and the camera-space normal cross into PyTorch for the field MLPs
    and the light (which stay PyTorch modules);
  * lerp-composite + dr.antialias per key (render.py:258-268) is one gather kernel per key, no composited image in HBM.
"""
import os

import torch

from .. import field_mlp
from .. import ops

# A/B switch for measurements: B2A_FUSE_AA_PAIR=0 renders the training pair of keys with two single-key launches
FUSE_MSAA_RESOLVE = os.environ.get("B2A_FUSE_MSAA_RESOLVE", "1") != "0"
FUSE_AA_PAIR = os.environ.get("B2A_FUSE_AA_PAIR", "1") != "0"


def interpolate(attr, rast, attr_idx, rast_db=None):
    """Reference render.py:23-24 helper (kept for callers such as render_uv); returns (out, None) like dr.interpolate."""
    return ops.interpolate(attr.contiguous(), rast, attr_idx), None


def _nearest_up(x, spp):
    return x.repeat_interleave(spp, dim=1).repeat_interleave(spp, dim=2) if spp > 1 else x


def _as_b3(x, B, device):
    x = torch.as_tensor(x, dtype=torch.float32, device=device)
    x = x.reshape(-1, 3)
    return x.expand(B, 3) if x.shape[0] == 1 and B > 1 else x


def _covered_rows(rast_s, rast_full=None, spp=1):
    """Flat indices (into [B*h*w]) of the shaded pixels that can reach an output, image index of each, and the grid shape.
    One device->host read (the row count) per render.  spp == 1: the pixels where a triangle is visible.  msaa (spp > 1,
    render.py:170-172,217-219): the shaded value of a low-resolution pixel is up-sampled to its spp x spp block and composited
    with the FULL-resolution alpha, so it is needed whenever ANY sub-pixel of the block is covered - also when the block's own
    nearest sample is not (the reference then shades material.sample(gb_tex_pos = 0), and that value carries gradient)."""
    B, h, w = rast_s.shape[:3]
    cov = rast_s[..., 3] > 0
    if rast_full is not None and spp > 1:
        cov = (rast_full[..., 3] > 0).view(B, h, spp, w, spp).any(4).any(2)
    idx = torch.nonzero(cov.reshape(-1)).squeeze(1)
    return idx, torch.div(idx, h * w, rounding_mode="floor"), (B, h, w)


# A/B switch for measurements: B2A_SPLIT_FIELD_FEAT=0 feeds the per-image feature to the field MLP as [N,C] rows (the
# reference's concat) instead of as a per-image bias of its first hidden layer
SPLIT_FIELD_FEAT = os.environ.get("B2A_SPLIT_FIELD_FEAT", "1") != "0"


def _splits_feat(net, feat):
    """True for a `CoordMLP` (MLPs.py:34-101, the reference's class or this package's twin) that takes a per-image feature."""
    if not SPLIT_FIELD_FEAT or feat is None or type(net).__name__ != "CoordMLP":
        return False
    layers = getattr(getattr(net, "mlp", None), "network", None)
    if layers is None or len(layers) == 0 or not hasattr(net, "in_layer"):
        return False
    first = layers[0]
    return (isinstance(first, torch.nn.Linear) and first.bias is None and getattr(net, "extra_feat_dim", 0) > 0
            and first.in_features == net.in_layer.out_features + net.extra_feat_dim and feat.shape[-1] == net.extra_feat_dim)


def _coord_mlp_rows(net, x, feat, img):
    """CoordMLP.forward (MLPs.py:72-98) on rows x [N,3] whose feature is feat[img[n]] ([B,C] per-image rows): the first
    hidden layer `Linear(nf + C -> nf)` of `relu(cat(h, feat))` (:90-94) is evaluated as W[:, :nf] . relu(h) + a PER-IMAGE
    bias W[:, nf:] . relu(feat_b) - the feature half of that layer (half of its FLOPs, 1 of the texture field's 8.3
    256x256-layer equivalents) runs on B rows instead of N, and neither the [N,C] feature rows nor the [N, nf+C] concat are
    ever materialised.  Same arithmetic up to the fp32 summation order of that one layer; autograd supplies the backward."""
    if net.symmetrize:
        x = torch.cat([x[..., :1].abs(), x[..., 1:]], -1)
    h = x
    if net.embedder is not None:
        h = net.embedder(x)
        if net.embed_concat_pts:
            h = torch.cat([x, h], -1)
    h = net.in_layer(h)
    layers = net.mlp.network
    w = layers[0].weight
    nf = net.in_layer.out_features
    bias = torch.nn.functional.linear(torch.relu(feat), w[:, nf:])             # [B, nf]
    z = torch.addmm(bias.index_select(0, img), torch.relu(h), w[:, :nf].t())    # `in_layer_relu` is idempotent under this relu
    out = layers[1:](z)
    if net.min_max is not None:
        out = out * (net.min_max[:, 1] - net.min_max[:, 0]) + net.min_max[:, 0]
    return out


def _sample_field(net, gb_tex_pos, feat, sparse):
    """net.sample on every pixel (sparse=None: the reference's evaluation) or on the covered rows only, scattered back
    into a zero image.  Same values on those rows (a per-image feature enters as a per-image bias, _coord_mlp_rows);
    the pixels left out never reach an output (alpha = 0 at every full-resolution sub-pixel of theirs in the composite,
    render.py:258-262)."""
    if sparse is None or getattr(net, "dense_only", False):
        return net.sample(gb_tex_pos, feat=feat)
    idx, img, (B, h, w) = sparse
    x = gb_tex_pos.reshape(-1, gb_tex_pos.shape[-1]).index_select(0, idx)
    if field_mlp.supported(net, x, feat):
        # CoordMLP on the tensor cores (csrc/field_mlp.cu): tcgen05 GEMMs over the covered rows, forward and backward
        y = field_mlp.coord_mlp_rows(net, x, None if feat is None else (feat if feat.shape[0] == B else feat.expand(B, -1)), img, B)
    elif _splits_feat(net, feat):
        y = _coord_mlp_rows(net, x, feat if feat.shape[0] == B else feat.expand(B, -1), img)
    else:
        f = None
        if feat is not None:
            f = feat.index_select(0, img) if feat.shape[0] == B else feat.expand(B, -1).index_select(0, img)
        y = net.sample(x, feat=f)
    if y.is_cuda and y.dtype == torch.float32:
        out = ops.scatter_rows(y, idx, B * h * w)
    else:           # autocast outputs in half precision (and the host-side tests of this glue): torch's own scatter
        out = y.new_zeros(B * h * w, y.shape[-1]).index_copy(0, idx, y)
    return out.view(B, h, w, y.shape[-1])


_AA_KEYS = ("shaded", "flow", "dino_pred", "depth", "shading")
_BG_KEYS = ("shaded", "geo_normal", "shading")
_KEEP = {"kd": 3, "ks": 3, "normal": 3, "geo_normal": 3, "shading": 1, "flow": 2, "depth": 1}


def render_mesh(ctx, mesh, mtx_in, w2c, view_pos, material, lgt, resolution, spp=1, num_layers=1, msaa=False, background=None,
                bsdf=None, feat=None, render_modes=None, prior_mesh=None, two_sided_shading=True, dino_net=None,
                num_frames=None, class_vector=None, sparse_fields=True):
    assert mesh.t_pos_idx.shape[1] > 0, "Got empty training triangle mesh (unrecoverable discontinuity)"
    assert background is None or (background.shape[1] == resolution[0] and background.shape[2] == resolution[1])
    if num_layers != 1:
        raise NotImplementedError("depth peeling beyond the first layer is never requested (AnimalModel.py:247)")
    if render_modes is None:
        render_modes = ["shaded"]
    dev = mesh.v_pos.device
    B = mesh.v_pos.shape[0]
    H, W = int(resolution[0]), int(resolution[1])
    shade_spp = spp if (spp > 1 and msaa) else 1          # g-buffer is shaded at [H,W] only when msaa is on
    full_res = (H * spp, W * spp)
    gH, gW = (H, W) if shade_spp > 1 else full_res
    mtx_in = torch.as_tensor(mtx_in, dtype=torch.float32, device=dev)
    campos = _as_b3(view_pos, B, dev)
    w2c = w2c.float()
    tri = mesh.tri_i32()
    opp = mesh.edge_adjacency()
    if prior_mesh is None:
        prior_mesh = mesh

    # clip-space transform + rasterize (render.py:278, :292-294) + fused g-buffer + antialias analysis: one autograd node
    want = ["cam_nrm", "tex_pos"]
    if "geo_normal" in render_modes:
        want.append("geo_nrm")
    if "normal" in render_modes:
        want.append("shading_nrm")
    if "depth" in render_modes:
        want.append("pos")
    need_aa = any(k in _AA_KEYS for k in render_modes)
    # msaa: rasterize at [H*spp, W*spp], shade the g-buffer at [H,W] from the nearest-downscaled rast (render.py:170-172);
    # otherwise the g-buffer lives at the raster resolution
    gres, gspp = ((H, W), spp) if shade_spp > 1 else (full_res, 1)
    v_pos_clip, rast, aa_ctx, gb = ops.render_geometry(mesh.v_pos, mesh.v_nrm, prior_mesh.v_pos, mtx_in, w2c, campos, tri, opp, gres,
                                                       spp=gspp, two_sided=two_sided_shading, want=tuple(want), need_aa=need_aa)
    rast_s = rast[:, ::shade_spp, ::shade_spp].contiguous() if shade_spp > 1 else rast
    gb_tex_pos, cam_normal = gb["tex_pos"], gb["cam_nrm"]

    # pixel shader: field MLPs + light stay PyTorch (render.py:50-94).  The fields are evaluated on COVERED pixels only
    # (SURVEY.md §8f-1): the reference runs both MLPs (1.6 MFLOP/pixel) on every pixel of the frame although the
    # composite discards everything where no triangle is visible - ~80 % of a 256^2 training view.
    nets = [n for n in (material, dino_net) if n is not None and not getattr(n, "dense_only", False)]
    sparse = _covered_rows(rast_s, rast, shade_spp) if (sparse_fields and nets) else None
    if material is not None:
        all_tex = _sample_field(material, gb_tex_pos, feat, sparse)
    else:
        all_tex = torch.ones(*gb_tex_pos.shape[:-1], 9, device=dev)
    kd, ks = all_tex[..., :3], all_tex[..., 3:6]
    bsdf = bsdf if bsdf is not None else getattr(material, "bsdf", None)
    assert bsdf is not None, "Material must specify a BSDF type"
    shading = None
    if bsdf == "diffuse":
        if lgt is None:
            shaded_col = kd
        elif type(lgt).__name__ == "EnvironmentLight":
            raise NotImplementedError("EnvironmentLight is used by no shipped config (SURVEY.md §2 #11)")
        else:
            shaded_col, shading = lgt.shade(feat, kd, cam_normal)
    else:
        raise NotImplementedError("bsdf '%s': only 'diffuse' is used by the shipped configs" % bsdf)
    buffers = {"shaded": shaded_col, "kd": kd, "ks": ks}
    if "geo_nrm" in gb:
        buffers["geo_normal"] = (gb["geo_nrm"] + 1.0) * 0.5
    if "shading_nrm" in gb:
        buffers["normal"] = (gb["shading_nrm"] + 1.0) * 0.5
    if shading is not None:
        buffers["shading"] = shading
    if dino_net is not None:
        buffers["dino_pred"] = _sample_field(dino_net, gb_tex_pos, class_vector, sparse)
    if "flow" in render_modes:  # render.py:281-288
        c2 = v_pos_clip[..., :2] / v_pos_clip[..., -1:]
        c2 = c2.view(-1, num_frames, *c2.shape[1:])
        dxy = c2[:, 1:] - c2[:, :-1]
        dxy = torch.cat([dxy, torch.zeros_like(dxy[:, :1])], dim=1).view(-1, *c2.shape[2:])
        buffers["flow"] = ops.interpolate(dxy, rast_s, tri)
    if "depth" in render_modes:  # render.py:102-108
        gb_pos = gb["pos"]
        hom = torch.cat([gb_pos, torch.ones_like(gb_pos[..., :1])], dim=-1)
        depth = torch.matmul(hom.view(B, -1, 4), w2c.transpose(-1, -2)).view(B, gH, gW, 4)[..., 2]
        dmin, dmax = depth.amin(dim=(1, 2), keepdim=True), depth.amax(dim=(1, 2), keepdim=True)
        buffers["depth"] = ((depth - dmin) / (dmax - dmin)).unsqueeze(-1)

    # background (render.py:298-304)
    if background is not None:
        bg_full = _nearest_up(background.float(), spp)
        bg_full = torch.cat((bg_full, torch.zeros_like(bg_full[..., 0:1])), dim=-1)
    else:
        bg_full = None

    # the training pair ['shaded', 'dino_pred'] goes through ONE composite+antialias launch per direction
    fused = {}
    if (FUSE_AA_PAIR and spp == 1 and aa_ctx is not None and "shaded" in render_modes and "dino_pred" in render_modes and "dino_pred" in buffers
            and buffers["shaded"].dtype == torch.float32 and buffers["dino_pred"].dtype == torch.float32
            and ops.pair_supported(buffers["dino_pred"], buffers["shaded"], aa_ctx)):
        dino_c = buffers["dino_pred"]
        fused["dino_pred"], fused["shaded"] = ops.composite_antialias_pair(dino_c, None, dino_c.shape[-1], buffers["shaded"], bg_full, 4,
                                                                           rast, v_pos_clip, tri, opp, aa_ctx, nchw=True)

    out_buffers = []
    for key in render_modes:
        if key not in buffers:
            out_buffers.append(None)
            continue
        if key in fused:
            out_buffers.append(fused[key])      # already the NCHW view
            continue
        color = buffers[key].float()
        bg = bg_full if key in _BG_KEYS else None
        if key == "shading" and bg is not None:
            bg = bg[..., 2:].contiguous()
        Cc = color.shape[-1] + 1
        keep = Cc if key == "shaded" else (Cc - 1 if key == "dino_pred" else _KEEP[key])
        if (shade_spp > 1 or key not in _AA_KEYS) and ops.composite_up_supported(color, aa_ctx):
            # msaa / logging keys: the low-resolution colour is up-sampled inside the composite kernel (no [B,H*spp,W*spp,C]
            # copies), un-antialiased keys composite in the same kernel instead of a torch lerp sequence
            pooled = FUSE_MSAA_RESOLVE and spp > 1 and shade_spp == spp
            accum = ops.composite_up(color, bg, v_pos_clip, full_res, up=shade_spp, antialias_edges=key in _AA_KEYS, keep=keep, aa_ctx=aa_ctx,
                                     pool=pooled)
            if pooled:      # the spp x spp average (render.py:322-323) came out of the same kernel
                out_buffers.append(accum.permute(0, 3, 1, 2))
                continue
        else:
            accum = ops.composite_antialias(_nearest_up(color, shade_spp), bg, rast, v_pos_clip, tri, opp, antialias_edges=key in _AA_KEYS,
                                            keep=keep, aa_ctx=aa_ctx)
        if spp > 1:
            accum = torch.nn.functional.avg_pool2d(accum.permute(0, 3, 1, 2), spp)
            out_buffers.append(accum)
        else:
            out_buffers.append(accum.permute(0, 3, 1, 2))
    return out_buffers


def render_uv(ctx, mesh, resolution, mlp_texture, feat=None):
    """Reference render.py:342-360: rasterize the UV atlas and sample the texture field at world positions."""
    uv_clip = mesh.v_tex * 2.0 - 1.0
    uv_clip4 = torch.cat((uv_clip, torch.zeros_like(uv_clip[..., 0:1]), torch.ones_like(uv_clip[..., 0:1])), dim=-1)
    rast = ops.rasterize(uv_clip4.contiguous(), mesh.t_tex_idx[0].int(), resolution)
    gb_pos, _ = interpolate(mesh.v_pos, rast, mesh.t_pos_idx[0].int())
    all_tex = mlp_texture.sample(gb_pos, feat=feat)
    assert all_tex.shape[-1] == 9 or all_tex.shape[-1] == 10, "Combined kd_ks_normal must be 9 or 10 channels"
    nrm = all_tex[..., -3:]
    nrm = nrm / torch.sqrt(torch.clamp(torch.sum(nrm * nrm, -1, keepdim=True), min=1e-20))
    return (rast[..., -1:] > 0).float(), all_tex[..., :-6], all_tex[..., -6:-3], nrm
