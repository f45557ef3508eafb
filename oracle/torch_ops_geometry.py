"""The reference's TORCH-OP geometry path restated device-agnostically (TEST / BASELINE INFRASTRUCTURE ONLY).

SURVEY.md §8d asks for a second baseline beside the CPU twin: what the reference's own formulation of R2 / R3 / R5 costs as
plain torch ops ON THE B200 ("the honest 'what the reference does today' for everything except the absent nvdiffrast").
The reference tree does not travel to the GPU box, so this module restates those three functions with the same torch
operations the reference issues (boolean-mask gathers, torch.unique(dim=0), scatter_add, per-bone matmuls), each citing
the file:line it follows.  Pinned on CPU against the reference-generated goldens (tests/test_oracle_golden.py); timed on
the GPU by bench.py's cpu_baseline leg.  Never imported by the product.
"""
import torch

from . import geometry_np as gnp


def marching_tets(pos, sdf, tets):
    """DMTet.__call__ (model/geometry/dmtet.py:104-155) in torch ops -> verts [V,3] (grad to sdf/pos), faces [F,3] i64."""
    dev = pos.device
    sdf = sdf.reshape(-1)
    with torch.no_grad():
        occ = sdf > 0                                                        # :106
        occ4 = occ[tets.reshape(-1)].reshape(-1, 4)                          # :107
        osum = occ4.sum(-1)                                                  # :108
        valid = (osum > 0) & (osum < 4)                                      # :109
        base = torch.as_tensor(gnp.BASE_TET_EDGES, device=dev)
        edges = tets[valid][:, base].reshape(-1, 2)                          # :113
        lo, hi = edges.min(1).values, edges.max(1).values                    # sort_edges :59-67
        uniq, inv = torch.unique(torch.stack([lo, hi], -1), dim=0, return_inverse=True)   # :115
        cross = occ[uniq.reshape(-1)].reshape(-1, 2).sum(-1) == 1            # :118
        mapping = torch.full((uniq.shape[0],), -1, dtype=torch.long, device=dev)          # :119
        mapping[cross] = torch.arange(int(cross.sum()), device=dev)          # :120
        idx_map = mapping[inv].reshape(-1, 6)                                # :121,133
        interp_v = uniq[cross]                                               # :123
    pa, pb = pos[interp_v[:, 0]], pos[interp_v[:, 1]]                        # :124-131
    sa, sb = sdf[interp_v[:, 0]], sdf[interp_v[:, 1]]
    verts = (pa * (-sb)[:, None] + pb * sa[:, None]) / (sa - sb)[:, None]
    with torch.no_grad():
        pow2 = torch.tensor([1, 2, 4, 8], device=dev)
        tetindex = (occ4[valid] * pow2).sum(-1)                              # :135-136
        tri_table = torch.as_tensor(gnp.TRIANGLE_TABLE, device=dev)
        num_tri = torch.as_tensor(gnp.NUM_TRIANGLES_TABLE, device=dev)[tetindex]          # :137
        m1, m2 = num_tri == 1, num_tri == 2
        f1 = torch.gather(idx_map[m1], 1, tri_table[tetindex[m1]][:, :3]).reshape(-1, 3)  # :141
        f2 = torch.gather(idx_map[m2], 1, tri_table[tetindex[m2]][:, :6]).reshape(-1, 3)  # :142
        faces = torch.cat([f1, f2], 0)
    return verts, faces


def auto_normals(v_pos, faces):
    """mesh.auto_normals (model/render/mesh.py:276-304): three scatter_adds with repeated index tensors."""
    B = v_pos.shape[0]
    i0, i1, i2 = faces[:, 0], faces[:, 1], faces[:, 2]
    v0, v1, v2 = v_pos[:, i0], v_pos[:, i1], v_pos[:, i2]
    fn = torch.cross(v1 - v0, v2 - v0, dim=-1)
    v_nrm = torch.zeros_like(v_pos)
    for idx in (i0, i1, i2):
        v_nrm = v_nrm.scatter_add(1, idx[None, :, None].repeat(B, 1, 3), fn)             # :291-295
    d = (v_nrm * v_nrm).sum(-1, keepdim=True)
    v_nrm = torch.where(d > 1e-20, v_nrm, torch.tensor([0.0, 0.0, 1.0], device=v_pos.device, dtype=v_pos.dtype))   # :297-298
    return v_nrm / torch.sqrt(torch.clamp((v_nrm * v_nrm).sum(-1, keepdim=True), min=1e-20))    # util.safe_normalize


def _affine(R, t):
    M = torch.zeros(R.shape[:-2] + (4, 4), device=R.device, dtype=R.dtype)
    M[..., :3, :3] = R
    M[..., :3, 3] = t
    M[..., 3, 3] = 1
    return M


def skinning(v_pos, bones, kinematic_tree, angles, temperature=1.0):
    """skinning (model/geometry/skinning.py:369-439): K-way stacked softmax weights on detached vertices, per-bone chain
    products Rest_i Rot(theta_i) Rest_i^-1 walked leaf -> root with one [B*F,V,4]@[4,4] matmul per bone (:399-431)."""
    B, F, K = angles.shape[:3]
    dev = v_pos.device
    bones = bones.expand(B, F, *bones.shape[2:])
    a, b = bones[:, :, :, 0, None, :], bones[:, :, :, 1, None, :]
    p = v_pos.detach()[:, :, None]                                            # :377
    ab = b - a
    t = ((p - a) * ab).sum(-1, keepdim=True) / torch.clamp((ab * ab).sum(-1, keepdim=True), min=1e-6)   # geometry/util.py:41-51
    d = torch.sqrt((((a + t.clamp(0.0, 1.0) * ab) - p) ** 2).sum(-1) + 1e-6)
    w = torch.softmax(-d / temperature, dim=2)                                # :16-22 [B,F,K,V]
    joint = bones[..., 0, :]
    fwd = torch.nn.functional.normalize(bones[..., 1, :] - joint, p=2, dim=-1)                          # :257
    right = torch.tensor([1.0, 0.0, 0.0], device=dev, dtype=fwd.dtype).expand_as(fwd)
    up = torch.nn.functional.normalize(torch.cross(fwd, right, dim=-1), p=2, dim=-1)                    # :261-262
    right = torch.cross(up, fwd, dim=-1)
    Rm = torch.stack([right, up, fwd], -1)                                    # :266
    rest = _affine(Rm, joint)
    rest_inv = _affine(Rm.transpose(-1, -2), -(Rm.transpose(-1, -2) @ joint[..., None])[..., 0])
    x, y, z = angles.unbind(-1)                                               # euler 'XYZ' :315-340
    cx, sx, cy, sy, cz, sz = x.cos(), x.sin(), y.cos(), y.sin(), z.cos(), z.sin()
    o, n = torch.ones_like(x), torch.zeros_like(x)
    Rx = torch.stack([o, n, n, n, cx, -sx, n, sx, cx], -1).reshape(x.shape + (3, 3))
    Ry = torch.stack([cy, n, sy, n, o, n, -sy, n, cy], -1).reshape(x.shape + (3, 3))
    Rz = torch.stack([cz, -sz, n, sz, cz, n, n, n, o], -1).reshape(x.shape + (3, 3))
    T = rest @ _affine(Rx @ Ry @ Rz, torch.zeros_like(joint)) @ rest_inv
    v4 = torch.cat([v_pos, torch.ones_like(v_pos[..., :1])], -1).expand(B, F, -1, -1)
    out = torch.zeros(B, F, v_pos.shape[2], 3, device=dev, dtype=v_pos.dtype)
    for k, chain in sorted(gnp.chain_lists(kinematic_tree).items()):
        M = T[:, :, chain[0]]
        for i in chain[1:]:                                                   # :399-417
            M = M @ T[:, :, i]
        xk = (v4 @ M.transpose(-1, -2))[..., :3]                              # :420-424, one matmul per bone
        out = out + w[:, :, k, :, None] * xk                                  # :428-431
    return out


def static_tables(tets, num_verts, tile, words):
    """The static per-grid tables in torch ops (TEST INFRASTRUCTURE: the checker of b2a_mt_build_edges / b2a_mt_build_tile_words):
    unique sorted (min,max) edges = the reference's generate_edges (dmtet.py:283-288) as a CSR by the smaller endpoint, and the
    tile skip table (distinct `vertex >> 5` words per tile of `tile` consecutive tets, ascending, padded with the smallest; slot 0
    = -1 when more than `words`).  -> edge_start [Vg+1] i32, edge_b [E] i32, tile_words [ceil(T/tile), words] i32."""
    tets = tets.long()
    T, n = tets.shape[0], int(num_verts) + 1
    be = torch.tensor([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], device=tets.device)
    e = tets[:, be].reshape(-1, 2)
    keys = torch.unique(e.min(1).values * n + e.max(1).values)
    start = torch.zeros(n, dtype=torch.int64, device=tets.device)
    start[1:] = torch.cumsum(torch.bincount(keys // n, minlength=n - 1), 0)
    nTT = (T + tile - 1) // tile
    w = (tets >> 5).reshape(-1)
    need = nTT * tile * 4
    if w.numel() < need:
        w = torch.cat([w, w[-4:].repeat((need - w.numel()) // 4)])
    ws = torch.sort(w.view(nTT, tile * 4), dim=1).values
    first = torch.ones_like(ws, dtype=torch.bool)
    first[:, 1:] = ws[:, 1:] != ws[:, :-1]
    rank = torch.cumsum(first, 1) - 1
    table = ws[:, :1].expand(-1, words).clone()
    sel = first & (rank < words)
    rows = torch.arange(nTT, device=tets.device)[:, None].expand_as(ws)[sel]
    table[rows, rank[sel]] = ws[sel]
    table[rank[:, -1] >= words, 0] = -1
    return start.int(), (keys % n).int(), table.int()
