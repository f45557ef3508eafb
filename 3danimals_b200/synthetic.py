"""Seeded synthetic inputs for the hot path: no tet grid, dataset or checkpoint
ships with the reference (data/tets/download_tets.sh needs network), so grids/SDFs/cameras are generated.

Pure numpy/torch-CPU helpers; they produce the reference's on-disk schema (npz with `vertices`, `indices`,
model/geometry/dmtet.py:223-225) so `DMTetGeometry.load_tets` keeps working.
"""
import itertools
import math
import os

import numpy as np


def kuhn_tet_grid(res):
    """(res+1)^3 lattice in [-0.5,0.5]^3, 6 Kuhn tets per cube (one per axis permutation).
    Returns vertices [Vg,3] f32, indices [T,4] i64 with Vg=(res+1)^3, T=6*res^3."""
    n = res + 1
    lin = np.linspace(-0.5, 0.5, n, dtype=np.float32)
    gx, gy, gz = np.meshgrid(lin, lin, lin, indexing="ij")
    verts = np.stack([gx, gy, gz], -1).reshape(-1, 3)
    ci, cj, ck = np.meshgrid(np.arange(res), np.arange(res), np.arange(res), indexing="ij")
    base = np.stack([ci, cj, ck], -1).reshape(-1, 3).astype(np.int64)
    vid = lambda p: (p[:, 0] * n + p[:, 1]) * n + p[:, 2]
    tets = []
    for perm in itertools.permutations(range(3)):
        p = base.copy()
        path = [vid(p)]
        for ax in perm:
            p = p.copy()
            p[:, ax] += 1
            path.append(vid(p))
        tets.append(np.stack(path, -1))
    tets = np.stack(tets, 1).reshape(-1, 4)
    # The six monotone paths alternate in orientation with the parity of the axis permutation.  Marching tetrahedra takes
    # its triangle winding from the tet's vertex order (triangle_table, dmtet.py:26-45), so a mixed grid yields a mesh
    # whose faces point in and out at random (area-weighted vertex normals then cancel).  Swap the last two vertices of
    # the positively oriented tets: every tet gets the same handedness and the extracted surface a consistent OUTWARD
    # winding, like the Quartet grids the reference downloads.
    p = verts[tets]
    vol = np.einsum("ni,ni->n", np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), p[:, 3] - p[:, 0])
    tets[vol > 0] = tets[vol > 0][:, [0, 1, 3, 2]]
    return verts, tets


def kuhn_tet_grid_torch(res, device):
    """kuhn_tet_grid on a torch device (identical tets in identical order; vertices equal up to linspace rounding): the
    res-256 grid has 100 M tets - minutes in numpy, a second on the GPU.  Returns vertices [Vg,3] f32, indices [T,4] i64."""
    import torch
    n = res + 1
    lin = torch.linspace(-0.5, 0.5, n, dtype=torch.float32, device=device)
    gx, gy, gz = torch.meshgrid(lin, lin, lin, indexing="ij")
    verts = torch.stack([gx, gy, gz], -1).reshape(-1, 3)
    r = torch.arange(res, device=device)
    ci, cj, ck = torch.meshgrid(r, r, r, indexing="ij")
    base = torch.stack([ci, cj, ck], -1).reshape(-1, 3)
    vid = lambda p: (p[:, 0] * n + p[:, 1]) * n + p[:, 2]
    tets = []
    for perm in itertools.permutations(range(3)):
        p = base.clone()
        path = [vid(p)]
        for ax in perm:
            p = p.clone()
            p[:, ax] += 1
            path.append(vid(p))
        # orientation = parity of the axis permutation (see kuhn_tet_grid): even permutations are the positively oriented ones
        parity = sum(1 for a in range(3) for b in range(a + 1, 3) if perm[a] > perm[b]) % 2
        if parity == 0:
            path[2], path[3] = path[3], path[2]
        tets.append(torch.stack(path, -1))
    return verts, torch.stack(tets, 1).reshape(-1, 4)


def write_tet_npz(res, root="data/tets"):
    """Writes data/tets/{res}_tets.npz in the reference schema (dmtet.py:223)."""
    os.makedirs(root, exist_ok=True)
    v, t = kuhn_tet_grid(res)
    path = os.path.join(root, "%d_tets.npz" % res)
    tmp = os.path.join(root, ".%d_tets.%d.tmp.npz" % (res, os.getpid()))
    np.savez(tmp, vertices=v, indices=t)
    os.replace(tmp, path)          # atomic: concurrent ranks see either no file or a complete one
    return path


def sdf_ellipsoid(pts, grid_scale=7.0):
    """The reference's 'ellipsoid' init SDF (dmtet.py:246-250), positive inside."""
    rxy = np.float32(grid_scale * 0.15)
    q = pts.astype(np.float32).copy()
    q[:, 2] = q[:, 2] / 2
    return (rxy - np.linalg.norm(q, axis=-1)).astype(np.float32)


def sdf_noisy_sphere(pts, radius=1.75, sigma=0.01, seed=0):
    rng = np.random.RandomState(seed)
    return (radius - np.linalg.norm(pts, axis=-1) + rng.randn(pts.shape[0]) * sigma).astype(np.float32)


def sdf_two_blobs(pts, r=1.0, sep=1.6):
    c0 = np.array([sep / 2, 0.2, 0.0], np.float32)
    c1 = np.array([-sep / 2, -0.2, 0.3], np.float32)
    d0 = r - np.linalg.norm(pts - c0, axis=-1)
    d1 = 0.8 * r - np.linalg.norm(pts - c1, axis=-1)
    return np.maximum(d0, d1).astype(np.float32)


def _capsule(pts, a, b, r):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    ab = b - a
    t = np.clip(((pts - a) @ ab) / max(float(ab @ ab), 1e-8), 0, 1)
    return r - np.linalg.norm(pts - (a + t[:, None] * ab), axis=-1)


def sdf_horse(pts, sigma=0.0, seed=0):
    """'Horse-like' union of capsules (body along z, 4 legs down -y, neck/head): makes the reference's
    quadrant leg detection (skinning.py:155-161) succeed. Positive inside. Extent ~[-0.6,0.6]x[-1.2,1.0]x[-1.6,1.9]."""
    pts = pts.astype(np.float32)
    parts = [
        _capsule(pts, [0, 0.15, -1.0], [0, 0.15, 1.0], 0.48),            # torso
        _capsule(pts, [0, 0.35, 1.0], [0, 0.95, 1.55], 0.27),            # neck
        _capsule(pts, [0, 0.95, 1.55], [0, 0.80, 1.95], 0.22),           # head
        _capsule(pts, [0.30, 0.0, 0.80], [0.33, -1.15, 0.85], 0.15),     # front right
        _capsule(pts, [-0.30, 0.0, 0.80], [-0.33, -1.15, 0.85], 0.15),   # front left
        _capsule(pts, [0.30, 0.0, -0.85], [0.33, -1.15, -0.95], 0.16),   # rear right
        _capsule(pts, [-0.30, 0.0, -0.85], [-0.33, -1.15, -0.95], 0.16), # rear left
        _capsule(pts, [0, 0.3, -1.1], [0, -0.2, -1.55], 0.08),           # tail
    ]
    s = np.max(np.stack(parts, 0), 0)
    if sigma > 0:
        s = s + np.random.RandomState(seed).randn(pts.shape[0]).astype(np.float32) * sigma
    return s.astype(np.float32)


def perspective(fovy=0.7854, aspect=1.0, n=0.1, f=1000.0):
    """render/util.py:189-194 (note the negated y row)."""
    y = np.tan(fovy / 2)
    return np.array([[1 / (y * aspect), 0, 0, 0],
                     [0, 1 / -y, 0, 0],
                     [0, 0, -(f + n) / (f - n), -(2 * f * n) / (f - n)],
                     [0, 0, -1, 0]], dtype=np.float32)


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32)


def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], np.float32)


def cameras(batch, seed=3, fov_deg=25.0, z_offset=10.0, znear=0.1, zfar=1000.0, max_az=math.pi, max_el=0.3,
            trans_sigma=0.05):
    """Seeded object poses -> (mvp [B,4,4], w2c [B,4,4], campos [B,3]) following
    InstancePredictorBase.get_camera_extrinsics_from_pose (InstancePredictorBase.py:606-620):
    w2c = [R | T + (0,0,-z_offset)], mvp = proj @ w2c, campos = -R^T T."""
    rng = np.random.RandomState(seed)
    proj = perspective(fov_deg / 180 * np.pi, 1.0, znear, zfar)
    mvp, w2c, campos = [], [], []
    for _ in range(batch):
        R = _rot_x(rng.uniform(-max_el, max_el)) @ _rot_y(rng.uniform(-max_az, max_az))
        T = (rng.randn(3) * trans_sigma).astype(np.float32) + np.array([0, 0, -z_offset], np.float32)
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = R
        m[:3, 3] = T
        w2c.append(m)
        mvp.append(proj @ m)
        campos.append(-(R.T @ T))
    return np.stack(mvp).astype(np.float32), np.stack(w2c).astype(np.float32), np.stack(campos).astype(np.float32)
