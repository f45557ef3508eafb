"""numpy restatement of the reference's geometry arithmetic (TEST INFRASTRUCTURE ONLY).

Every function cites the reference file:line it follows (paths relative to /root/reference).
Pinned against the reference's own Python by tests/test_oracle_vs_reference.py (build
container) and by the committed goldens under tests/golden/ (everywhere).
"""
import math

import numpy as np

# --- tables: data copied from model/geometry/dmtet.py:26-46 ("data, not code") ---
TRIANGLE_TABLE = np.array([
    [-1, -1, -1, -1, -1, -1], [1, 0, 2, -1, -1, -1], [4, 0, 3, -1, -1, -1], [1, 4, 2, 1, 3, 4],
    [3, 1, 5, -1, -1, -1], [2, 3, 0, 2, 5, 3], [1, 4, 0, 1, 5, 4], [4, 2, 5, -1, -1, -1],
    [4, 5, 2, -1, -1, -1], [4, 1, 0, 4, 5, 1], [3, 2, 0, 3, 5, 2], [1, 3, 5, -1, -1, -1],
    [4, 1, 2, 4, 3, 1], [3, 0, 4, -1, -1, -1], [2, 0, 1, -1, -1, -1], [-1, -1, -1, -1, -1, -1]],
    dtype=np.int64)
NUM_TRIANGLES_TABLE = np.array([0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0], dtype=np.int64)
BASE_TET_EDGES = np.array([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], dtype=np.int64)


def marching_tets(pos_nx3, sdf_n, tet_fx4, with_uvs=True):
    """DMTet.__call__ (model/geometry/dmtet.py:104-155).

    Returns dict(verts [V,3] f32, faces [F,3] i64, uv_idx [F,3] i64, uvs [4N^2,2] f32 (optional),
    interp_v [V,2] i64 = the (min,max) grid-vertex pair each output vertex lies on, face_gidx [F]).
    """
    pos = np.asarray(pos_nx3, dtype=np.float32)
    sdf = np.asarray(sdf_n, dtype=np.float32).reshape(-1)
    tet = np.asarray(tet_fx4, dtype=np.int64)
    occ_n = sdf > 0                                              # :106
    occ_fx4 = occ_n[tet.reshape(-1)].reshape(-1, 4)              # :107
    occ_sum = occ_fx4.sum(-1)                                    # :108
    valid = (occ_sum > 0) & (occ_sum < 4)                        # :109
    vt = tet[valid]
    all_edges = vt[:, BASE_TET_EDGES].reshape(-1, 2)             # :113
    all_edges = np.sort(all_edges, axis=1)                       # sort_edges :59-67
    if all_edges.shape[0] == 0:
        unique_edges = np.zeros((0, 2), np.int64)
        idx_map = np.zeros((0,), np.int64)
    else:
        # torch.unique(dim=0) == lexicographic order of rows (:115)
        key = all_edges[:, 0] * (sdf.shape[0] + 1) + all_edges[:, 1]
        ukey, idx_map = np.unique(key, return_inverse=True)
        unique_edges = np.stack([ukey // (sdf.shape[0] + 1), ukey % (sdf.shape[0] + 1)], -1)
    mask_edges = occ_n[unique_edges.reshape(-1)].reshape(-1, 2).sum(-1) == 1   # :118
    mapping = -np.ones(unique_edges.shape[0], np.int64)           # :119
    mapping[mask_edges] = np.arange(int(mask_edges.sum()))        # :120
    idx_map = mapping[idx_map].reshape(-1, 6)                      # :121,133
    interp_v = unique_edges[mask_edges]                            # :123
    verts = lerp_vertices(pos, sdf, interp_v)
    v_id = 2 ** np.arange(4, dtype=np.int64)                       # :135
    tetindex = (occ_fx4[valid] * v_id[None]).sum(-1)               # :136
    num_tri = NUM_TRIANGLES_TABLE[tetindex]                        # :137
    m1, m2 = num_tri == 1, num_tri == 2
    f1 = np.take_along_axis(idx_map[m1], TRIANGLE_TABLE[tetindex[m1]][:, :3], 1).reshape(-1, 3)  # :141
    f2 = np.take_along_axis(idx_map[m2], TRIANGLE_TABLE[tetindex[m2]][:, :6], 1).reshape(-1, 3)  # :142
    faces = np.concatenate([f1, f2], 0)
    num_tets = tet.shape[0]
    tet_gidx = np.arange(num_tets, dtype=np.int64)[valid]          # :147
    g2 = tet_gidx[m2] * 2
    face_gidx = np.concatenate([tet_gidx[m1] * 2, np.stack([g2, g2 + 1], -1).reshape(-1)], 0)   # :148-151
    out = dict(verts=verts, faces=faces, interp_v=interp_v, face_gidx=face_gidx)
    uvs, uv_idx = map_uv(face_gidx, num_tets * 2, with_uvs)
    out["uv_idx"] = uv_idx
    if with_uvs:
        out["uvs"] = uvs
    return out


def lerp_vertices(pos, sdf, interp_v):
    """Crossing-point interpolation, dmtet.py:124-131:  v = (p_a*(-s_b) + p_b*s_a) / (s_a - s_b),
    evaluated as p_a*((-s_b)/den) + p_b*(s_a/den) with every op rounded to fp32 (no FMA)."""
    a, b = interp_v[:, 0], interp_v[:, 1]
    sa = sdf[a].astype(np.float32)
    sb = (-sdf[b]).astype(np.float32)            # edges_to_interp_sdf[:,-1] *= -1  (:126)
    den = (sa + sb).astype(np.float32)           # :128
    wa = (sb / den).astype(np.float32)           # flip (:130): weight of p_a is (-s_b)/den
    wb = (sa / den).astype(np.float32)
    return ((pos[a] * wa[:, None]).astype(np.float32) + (pos[b] * wb[:, None]).astype(np.float32)).astype(np.float32)


def lerp_vertices_bwd(pos, sdf, interp_v, d_verts):
    """Analytic adjoint of lerp_vertices w.r.t. sdf (and pos); what autograd yields for dmtet.py:124-131."""
    a, b = interp_v[:, 0], interp_v[:, 1]
    sa = sdf[a].astype(np.float64)
    sb = sdf[b].astype(np.float64)
    den = sa - sb
    g = (d_verts.astype(np.float64) * (pos[a].astype(np.float64) - pos[b].astype(np.float64))).sum(-1)
    d_sdf = np.zeros(sdf.shape[0], np.float64)
    np.add.at(d_sdf, a, g * sb / den ** 2)
    np.add.at(d_sdf, b, -g * sa / den ** 2)
    d_pos = np.zeros(pos.shape, np.float64)
    np.add.at(d_pos, a, d_verts * (-sb / den)[:, None])
    np.add.at(d_pos, b, d_verts * (sa / den)[:, None])
    return d_sdf.astype(np.float32), d_pos.astype(np.float32)


def map_uv(face_gidx, max_idx, with_uvs=True):
    """DMTet.map_uv (dmtet.py:69-98)."""
    N = int(np.ceil(np.sqrt((max_idx + 1) // 2)))
    uvs = None
    if with_uvs:
        lin = np.linspace(0, 1 - (1 / N), N, dtype=np.float32)
        tex_y, tex_x = np.meshgrid(lin, lin, indexing="ij")
        pad = np.float32(0.9 / N)
        uvs = np.stack([tex_x, tex_y, tex_x + pad, tex_y, tex_x + pad, tex_y + pad, tex_x, tex_y + pad], -1)
        uvs = uvs.reshape(-1, 2).astype(np.float32)
    t = face_gidx // 2
    tet_idx = (t // N) * N + (t % N)
    tri_idx = face_gidx % 2
    uv_idx = np.stack([tet_idx * 4, tet_idx * 4 + tri_idx + 1, tet_idx * 4 + tri_idx + 2], -1).reshape(-1, 3)
    return uvs, uv_idx.astype(np.int64)


def unique_sorted_edges(tet_fx4):
    """DMTetGeometry.generate_edges (dmtet.py:283-288): unique (min,max) edges, lexicographic."""
    tet = np.asarray(tet_fx4, dtype=np.int64)
    e = np.sort(tet[:, BASE_TET_EDGES].reshape(-1, 2), axis=1)
    n = int(tet.max()) + 2
    key = np.unique(e[:, 0] * n + e[:, 1])
    return np.stack([key // n, key % n], -1)


# ---------------------------------------------------------------------------------------------
# vertex normals  (model/render/mesh.py:276-304)
# ---------------------------------------------------------------------------------------------
def auto_normals(v_pos, faces):
    """mesh.auto_normals: area-weighted face-normal splat, fallback (0,0,1), safe_normalize (util.py:28-32)."""
    v_pos = np.asarray(v_pos, np.float32)
    i0, i1, i2 = faces[:, 0], faces[:, 1], faces[:, 2]
    v0, v1, v2 = v_pos[:, i0], v_pos[:, i1], v_pos[:, i2]
    fn = np.cross(v1 - v0, v2 - v0).astype(np.float32)
    nsum = np.zeros_like(v_pos)
    for b in range(v_pos.shape[0]):
        for idx in (i0, i1, i2):
            np.add.at(nsum[b], idx, fn[b])
    d = (nsum * nsum).sum(-1, keepdims=True)
    nsum = np.where(d > 1e-20, nsum, np.array([0, 0, 1], np.float32))
    d = (nsum * nsum).sum(-1, keepdims=True)
    return (nsum / np.sqrt(np.maximum(d, 1e-20))).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# bones + LBS  (model/geometry/skinning.py, model/geometry/util.py)
# ---------------------------------------------------------------------------------------------
def line_segment_distance(a, b, points):
    """geometry/util.py:30-53.  a,b [...,3]; points [...,V,3] -> [...,V]."""
    a = a[..., None, :]
    b = b[..., None, :]
    ab = b - a
    t = ((points - a) * ab).sum(-1, keepdims=True) / np.maximum((ab * ab).sum(-1, keepdims=True), np.float32(1e-6))
    t = np.clip(t, 0.0, 1.0)
    s = a + t * ab
    return np.sqrt(((s - points) ** 2).sum(-1) + np.float32(1e-6))


def skinning_weights(bones, v_pos, temperature):
    """_compute_vertices_to_bones_weights (skinning.py:16-22): softmax_k(-dist/T) -> [K,B,F,V]."""
    K = bones.shape[2]
    d = np.stack([line_segment_distance(bones[:, :, k, 0], bones[:, :, k, 1], v_pos) for k in range(K)], 0)
    x = -d / temperature
    x = x - x.max(0, keepdims=True)
    e = np.exp(x)
    return (e / e.sum(0, keepdims=True)).astype(np.float32)


def euler_xyz(angles):
    """euler_angles_to_matrix(..., 'XYZ') = Rx @ Ry @ Rz (skinning.py:289-340)."""
    x, y, z = angles[..., 0], angles[..., 1], angles[..., 2]
    cx, sx, cy, sy, cz, sz = np.cos(x), np.sin(x), np.cos(y), np.sin(y), np.cos(z), np.sin(z)
    one, zero = np.ones_like(x), np.zeros_like(x)
    Rx = np.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], -1).reshape(x.shape + (3, 3))
    Ry = np.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], -1).reshape(x.shape + (3, 3))
    Rz = np.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], -1).reshape(x.shape + (3, 3))
    return Rx @ Ry @ Rz


def bone_rest_frames(bones):
    """_estimate_bone_rotation (skinning.py:251-270): columns [right | up | forward]; + joint translation."""
    joint = bones[..., 0, :]
    fwd = bones[..., 1, :] - bones[..., 0, :]
    fwd = fwd / np.maximum(np.linalg.norm(fwd, axis=-1, keepdims=True), 1e-12)
    right0 = np.broadcast_to(np.array([1, 0, 0], fwd.dtype), fwd.shape)
    up = np.cross(fwd, right0)
    up = up / np.maximum(np.linalg.norm(up, axis=-1, keepdims=True), 1e-12)
    right = np.cross(up, fwd)
    up = up / np.maximum(np.linalg.norm(up, axis=-1, keepdims=True), 1e-12)
    R = np.stack([right, up, fwd], -1)
    M = np.zeros(bones.shape[:-2] + (4, 4), bones.dtype)
    M[..., :3, :3] = R
    M[..., :3, 3] = joint
    M[..., 3, 3] = 1
    Minv = np.zeros_like(M)
    Rt = np.swapaxes(R, -1, -2)
    Minv[..., :3, :3] = Rt
    Minv[..., :3, 3] = -(Rt @ joint[..., None])[..., 0]
    Minv[..., 3, 3] = 1
    return M, Minv


def chain_lists(kinematic_tree):
    """Ancestor lists in application order (root ... parent, bone): skinning.py:389-396."""
    out = {}
    for bone_id, _ in kinematic_tree:
        parents = [p for p, children in kinematic_tree if bone_id in children]
        out[bone_id] = parents + [bone_id]
    return out


def bone_transforms(bones, angles, kinematic_tree):
    """Per-bone global 4x4 (skinning.py:398-417): G_k = T_root ... T_parent T_k, T_i = Rest_i Rot(theta_i) Rest_i^-1.
    bones [Bb,Fb,K,2,3] (Bb,Fb broadcastable), angles [B,F,K,3] -> G [B,F,K,4,4]."""
    B, F, K = angles.shape[:3]
    bones = np.broadcast_to(bones, (B, F) + bones.shape[2:]).astype(np.float64)
    Rest, RestInv = bone_rest_frames(bones)
    Rot = np.zeros((B, F, K, 4, 4))
    Rot[..., :3, :3] = euler_xyz(angles.astype(np.float64))
    Rot[..., 3, 3] = 1
    T = Rest @ Rot @ RestInv
    G = np.zeros((B, F, K, 4, 4))
    for k, chain in chain_lists(kinematic_tree).items():
        M = np.broadcast_to(np.eye(4), (B, F, 4, 4)).copy()
        for i in chain:           # root first: M = T_root @ ... @ T_k
            M = M @ T[:, :, i]
        G[:, :, k] = M
    return G


def skinning(v_pos, bones, kinematic_tree, angles, temperature=1.0):
    """skinning (skinning.py:369-439).  v_pos [Bv,Fv,V,3], bones [Bb,Fb,K,2,3], angles [B,F,K,3].
    Returns verts [B,F,V,3], weights [K,Bv',Fv',V], posed_bones [B,F,K,2,3] (fp64 accumulate, cast fp32)."""
    B, F, K = angles.shape[:3]
    w = skinning_weights(bones.astype(np.float32), v_pos.astype(np.float32), np.float32(temperature)).astype(np.float64)
    G = bone_transforms(bones, angles, kinematic_tree)
    vp = np.broadcast_to(v_pos, (B, F) + v_pos.shape[2:]).astype(np.float64)
    v4 = np.concatenate([vp, np.ones(vp.shape[:-1] + (1,))], -1)
    out = np.zeros(vp.shape)
    for k in range(K):
        xk = np.einsum("bfij,bfvj->bfvi", G[:, :, k], v4)[..., :3]
        out += w[k][..., None] * xk
    b4 = np.concatenate([np.broadcast_to(bones, (B, F) + bones.shape[2:]).astype(np.float64),
                         np.ones((B, F, K, 2, 1))], -1)
    posed = np.einsum("bfkij,bfkej->bfkei", G, b4)[..., :3]
    return out.astype(np.float32), w.astype(np.float32), posed.astype(np.float32)


def estimate_bones(seq_shape, n_body_bones, n_legs=4, n_leg_bones=0, body_bones_mode="z_minmax",
                   compute_kinematic_chain=True, aux=None, attach_legs_to_body=True,
                   legs_to_body_joint_indices=None, bone_y_threshold=None):
    """estimate_bones (skinning.py:49-248), numpy restatement. seq_shape [B,F,V,3]."""
    s = np.asarray(seq_shape, np.float32)
    B, F, V, _ = s.shape
    bi, fi = np.meshgrid(np.arange(B), np.arange(F), indexing="ij")
    if body_bones_mode == "z_minmax":                               # :69-75
        pa = s[bi, fi, s[..., 2].argmax(2)].copy()
        pb = s[bi, fi, s[..., 2].argmin(2)].copy()
    elif body_bones_mode == "z_minmax_y+":                          # :76-87
        mid = s.mean(2)
        m = (s[..., 1] > (mid[:, :, None, 1] - 0.5)).astype(np.float32)
        pa = s[bi, fi, (s[..., 2] * m + np.float32(-1e6) * (1 - m)).argmax(2)].copy()
        pb = s[bi, fi, (s[..., 2] * m + np.float32(1e6) * (1 - m)).argmin(2)].copy()
    else:
        raise NotImplementedError
    pa[..., 0] = 0                                                   # :91-92
    pb[..., 0] = 0
    mid = s.mean(2, dtype=np.float32)
    mid[..., 0] = 0
    if n_leg_bones > 0:
        mid[..., 1] += 0.5                                           # :98-99
    assert n_body_bones % 2 == 0
    n_joints = n_body_bones + 1
    blend = np.linspace(0., 1., math.ceil(n_joints / 2), dtype=np.float32)[None, None, :, None]
    ja = pa[:, :, None] * (1 - blend) + mid[:, :, None] * blend      # :104
    jb = pb[:, :, None] * blend + mid[:, :, None] * (1 - blend)      # :106
    joints = np.concatenate([ja[:, :, :-1], jb], 2)
    if compute_kinematic_chain:                                      # :111-131
        aux = {}
        half = n_body_bones // 2
        b2j, chain, bone_idx, dep = [], [], 0, []
        for i in range(half):
            b2j.append((i + 1, i))
            chain = [(bone_idx, dep)] + chain
            dep = dep + [bone_idx]
            bone_idx += 1
        dep = []
        for i in range(n_body_bones - 1, half - 1, -1):
            b2j.append((i, i + 1))
            chain = [(bone_idx, dep)] + chain
            dep = dep + [bone_idx]
            bone_idx += 1
        aux["bones_to_joints"] = b2j
    else:
        b2j, chain = aux["bones_to_joints"], aux["kinematic_chain"]
    j2b = lambda J, idx: np.stack([np.stack([J[:, :, a], J[:, :, b]], 2) for a, b in idx], 2)   # :8-13
    bones_pred = j2b(joints, b2j)
    if n_leg_bones > 0:
        assert n_legs == 4
        xs, ys, zs = s[..., 0], s[..., 1], s[..., 2]
        if bone_y_threshold is None:                                 # :156-161
            x_margin = (np.quantile(xs, 0.95) - np.quantile(xs, 0.05)).astype(np.float32) * np.float32(0.2)
            quads = [(xs > x_margin) & (zs > 0), (xs > x_margin) & (zs < 0),
                     (xs < -x_margin) & (zs < 0), (xs < -x_margin) & (zs > 0)]
        else:                                                        # :163-175
            flags = ys < np.quantile(ys, bone_y_threshold)
            x0, z0 = np.quantile(xs[flags], 0.5), np.quantile(zs[flags], 0.5)
            xm = (np.quantile(xs[flags], 0.95) - np.quantile(xs[flags], 0.05)) * 0.2
            zm = (np.quantile(zs[flags], 0.95) - np.quantile(zs[flags], 0.05)) * 0.2
            quads = [(xs - x0 > xm) & (zs - z0 > zm), (xs - x0 > xm) & (zs < z0),
                     (xs - x0 < -xm) & (zs < z0), (xs - x0 < -xm) & (zs - z0 > zm)]

        def find_leg(quad, body_bone_idx):                           # :177-198
            allj = np.zeros((B, F, n_leg_bones + 1, 3), np.float32)
            for b in range(B):
                for f in range(F):
                    pts = s[b, f][quad[b, f]]
                    foot = pts[np.argmin(pts[:, 1])]
                    if body_bone_idx is None:
                        body_bone_idx = int(np.argmin(np.abs(bones_pred[b, f, :, 1, 2] - foot[2])))
                    bj = bones_pred[b, f, body_bone_idx, 1]
                    bl = np.linspace(0., 1., n_leg_bones + 1, dtype=np.float32)[:, None]
                    allj[b, f] = foot[None] * (1 - bl) + bj[None] * bl
            return allj, body_bone_idx

        if legs_to_body_joint_indices is None:
            legs_to_body_joint_indices = [None] * 4
        start = n_body_bones
        leg_bones_all = []
        leg_auxs = [] if compute_kinematic_chain else aux["legs"]
        for i, quad in enumerate(quads):
            if compute_kinematic_chain:                              # :210-226
                bb = legs_to_body_joint_indices[i]
                if i == 2:
                    bb = legs_to_body_joint_indices[1]
                elif i == 3:
                    bb = legs_to_body_joint_indices[0]
                lj, bb = find_leg(quad, bb)
                legs_to_body_joint_indices[i] = bb
                lb2j, lchain, lidx, ldep, bidx = [], [], [], [], start      # build_kinematic_chain :25-37
                for j in range(n_leg_bones):
                    lb2j.append((j + 1, j))
                    lchain = [(bidx, ldep)] + lchain
                    ldep = ldep + [bidx]
                    bidx += 1
                lidx = ldep
                if attach_legs_to_body:                              # update_body_kinematic_chain :40-46
                    for bone_id, deps in chain:
                        if bone_id == bb or bb in deps:
                            deps += lidx
                chain = chain + lchain
                leg_auxs.append(dict(body_bone_idx=bb, leg_bones_to_joints=lb2j))
                start += n_leg_bones
            else:
                bb = leg_auxs[i]["body_bone_idx"]
                lj, _ = find_leg(quad, bb)
                lb2j = leg_auxs[i]["leg_bones_to_joints"]
            leg_bones_all.append(j2b(lj, lb2j))
        all_bones = np.concatenate([bones_pred] + leg_bones_all, 2)
    else:
        all_bones = bones_pred
    if compute_kinematic_chain:
        aux["kinematic_chain"] = chain
        if n_leg_bones > 0:
            aux["legs"] = leg_auxs
        return all_bones, chain, aux
    return all_bones
