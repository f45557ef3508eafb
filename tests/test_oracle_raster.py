"""Self-consistency of the C restatement of the nvdiffrast ops (oracle/raster_ref.c).  No reference golden exists at
this boundary (parity unpinned), so the restatement is held to analytic answers and to finite differences
of its own forward passes."""
import numpy as np

from oracle import raster as R


def test_single_triangle_coverage_and_barycentrics():
    pos = np.array([[[-1, -1, 0.25, 1], [1, -1, 0.25, 1], [-1, 1, 0.25, 1]]], np.float32)
    tri = np.array([[0, 1, 2]], np.int32)
    H = W = 8
    rast = R.rasterize(pos, tri, (H, W))
    for py in range(H):
        for px in range(W):
            fx, fy = (px + 0.5) / W * 2 - 1, (py + 0.5) / H * 2 - 1          # row 0 is clip y = -1
            inside = fx + fy <= 1e-6
            assert (rast[0, py, px, 3] == 1.0) == inside, (px, py)
            if inside and fx + fy < -1e-3:
                u, v = rast[0, py, px, 0], rast[0, py, px, 1]
                # weights of vertices 0 and 1: p = u*v0 + v*v1 + (1-u-v)*v2
                assert abs((u * -1 + v * 1 + (1 - u - v) * -1) - fx) < 1e-5
                assert abs((u * -1 + v * -1 + (1 - u - v) * 1) - fy) < 1e-5
                assert abs(rast[0, py, px, 2] - 0.25) < 1e-6
    assert np.all(rast[rast[..., 3] == 0] == 0)


def test_depth_order_tie_break_and_clipping():
    quad = lambda z: [[-0.9, -0.9, z, 1], [0.9, -0.9, z, 1], [0.9, 0.9, z, 1], [-0.9, 0.9, z, 1]]
    pos = np.array([quad(0.5) + quad(0.2) + quad(0.2)], np.float32)
    tri = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [8, 9, 10], [8, 10, 11]], np.int32)
    ids = R.rasterize(pos, tri, (16, 16))[0, ..., 3]
    assert set(np.unique(ids)) <= {0.0, 3.0, 4.0}            # nearer quad wins; equal depth -> lowest triangle index
    # depth outside [-1,1] and geometry behind the eye are rejected
    far = np.array([[[-1, -1, 1.5, 1], [1, -1, 1.5, 1], [0, 1, 1.5, 1], [-1, -1, 0, -1], [1, -1, 0, -1], [0, 1, 0, -1]]], np.float32)
    assert np.all(R.rasterize(far, np.array([[0, 1, 2], [3, 4, 5]], np.int32), (8, 8))[..., 3] == 0)
    # perspective-correct barycentrics: w varies across the triangle
    p = np.array([[[-2, -2, 0.2, 2.0], [3, -3, 0.9, 3.0], [0, 1, 0.1, 1.0]]], np.float32)
    r = R.rasterize(p, np.array([[0, 1, 2]], np.int32), (32, 32))
    m = r[0, ..., 3] > 0
    u, v = r[0, ..., 0][m], r[0, ..., 1][m]
    clip = u[:, None] * p[0, 0] + v[:, None] * p[0, 1] + (1 - u - v)[:, None] * p[0, 2]
    ys, xs = np.nonzero(m)
    assert np.allclose(clip[:, 0] / clip[:, 3], (xs + 0.5) / 32 * 2 - 1, atol=2e-5)
    assert np.allclose(clip[:, 1] / clip[:, 3], (ys + 0.5) / 32 * 2 - 1, atol=2e-5)
    assert np.allclose(clip[:, 2] / clip[:, 3], r[0, ..., 2][m], atol=2e-5)


def edge_tie_scene(res=8):
    """Two coplanar triangles (a quad split along its diagonal) whose shared edge and outer edges pass EXACTLY through pixel
    centres, drawn twice at the same depth: every tie the fill rule has to break."""
    s = 2.0 / res                                   # one pixel in NDC; pixel centres at -1 + (i + .5) s
    lo, hi = -1 + 1.5 * s, -1 + (res - 1.5) * s     # the quad's corners sit ON pixel centres
    quad = [[lo, lo, 0.25, 1], [hi, lo, 0.25, 1], [hi, hi, 0.25, 1], [lo, hi, 0.25, 1]]
    pos = np.array([quad + quad], np.float32)
    tri = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.int32)
    return pos, tri


def test_fill_rule_on_exact_edge_and_depth_ties():
    """Documents where this restatement's fill rule can differ from nvdiffrast's OpenGL rasterizer (top-left rule, draw order):
    a pixel centre exactly on an edge is INSIDE (inclusive edges), a centre on the shared diagonal is claimed by both triangles and
    goes to the lower triangle index, equal depth goes to the lower index.  GL's top-left rule would drop the centres on the
    right / bottom outer edges; on meshes in general position (every test outside this one) no centre lies exactly on an edge."""
    res = 8
    pos, tri = edge_tie_scene(res)
    ids = R.rasterize(pos, tri, (res, res))[0, ..., 3]
    inside = np.zeros((res, res), bool)
    inside[1:res - 1, 1:res - 1] = True              # centres 1..res-2 in x and y: the closed quad, outer edges included
    assert np.array_equal(ids > 0, inside)
    ys, xs = np.nonzero(inside)
    # triangle 0 = (lo,lo),(hi,lo),(hi,hi): the half with x >= y; the diagonal x == y is shared -> lowest index (0 -> id 1)
    want = np.where(xs >= ys, 1.0, 2.0)
    assert np.array_equal(ids[inside], want)
    assert not np.isin(ids, [3.0, 4.0]).any()        # the second copy at equal depth never wins


def _fd(f, x, g, idx, eps):
    """directional finite difference of sum(f(x) * g) along coordinate idx (float64 accumulation)"""
    xp, xm = x.copy(), x.copy()
    xp[idx] += eps
    xm[idx] -= eps
    return float(((f(xp).astype(np.float64) - f(xm).astype(np.float64)) * g).sum() / (2 * eps))


def _small_scene():
    rng = np.random.RandomState(0)
    pos = np.array([[[-0.7, -0.6, 0.3, 1.0], [0.8, -0.5, 0.4, 1.2], [0.1, 0.75, 0.2, 0.9], [0.9, 0.8, 0.6, 1.1], [-0.8, 0.7, 0.5, 1.0]]], np.float32)
    tri = np.array([[0, 1, 2], [1, 3, 2], [0, 2, 4]], np.int32)
    return rng, pos, tri


def test_interpolate_and_rasterize_backward_match_finite_differences():
    rng, pos, tri = _small_scene()
    res = (24, 24)
    rast = R.rasterize(pos, tri, res)
    attr = rng.randn(1, 5, 3).astype(np.float32)
    g = rng.randn(1, 24, 24, 3)
    da, dr = R.interpolate_bwd(attr, rast, tri, g.astype(np.float32))
    for idx in [(0, 0, 0), (0, 2, 1), (0, 4, 2)]:
        assert abs(_fd(lambda a: R.interpolate(a, rast, tri), attr, g, idx, 1e-2) - da[idx]) < 2e-3 * max(1, abs(da[idx]))
    # barycentric gradient -> clip positions: differentiate interpolate(rasterize(pos)) with triangle ids held fixed,
    # i.e. only on pixels whose id does not change under the perturbation
    d_pos = R.rasterize_bwd(pos, tri, rast, dr)
    for idx in [(0, 0, 0), (0, 1, 1), (0, 2, 3), (0, 3, 0)]:
        eps = 1e-3
        xp, xm = pos.copy(), pos.copy()
        xp[idx] += eps
        xm[idx] -= eps
        rp, rm = R.rasterize(xp, tri, res), R.rasterize(xm, tri, res)
        keep = (rp[..., 3] == rast[..., 3]) & (rm[..., 3] == rast[..., 3]) & (rast[..., 3] > 0)
        # keep only pixels strictly inside (unsaturated barycentrics) for a clean derivative
        fd = (((R.interpolate(attr, rp, tri).astype(np.float64) - R.interpolate(attr, rm, tri)) * g).sum(-1) * keep).sum() / (2 * eps)
        ref = R.rasterize_bwd(pos, tri, rast, dr * keep[..., None])[idx]
        assert abs(fd - ref) < 2e-2 * max(1.0, abs(ref)), (idx, fd, ref)
    assert np.all(d_pos[..., 2] == 0)        # nothing flows through z (only x, y, w)


def test_antialias_blends_silhouette_only_and_backward_matches_finite_differences():
    rng, pos, tri = _small_scene()
    res = (24, 24)
    rast = R.rasterize(pos, tri, res)
    opp = R.edge_adjacency(tri, 5)
    assert opp.tolist() == [[3, 4, -1], [-1, 0, -1], [-1, -1, 1]]
    color = rng.rand(1, 24, 24, 3).astype(np.float32) * (rast[..., 3:] > 0) + 0.1
    out = R.antialias(color, rast, pos, tri, opp)
    changed = np.abs(out - color).max(-1)[0] > 0
    ids = rast[0, ..., 3]
    # only pixels next to an id change can be touched, and interior shared edges (0-1-2 / 1-3-2) are not silhouettes
    nb = np.zeros_like(changed)
    nb[:, 1:] |= ids[:, 1:] != ids[:, :-1]
    nb[:, :-1] |= ids[:, 1:] != ids[:, :-1]
    nb[1:, :] |= ids[1:, :] != ids[:-1, :]
    nb[:-1, :] |= ids[1:, :] != ids[:-1, :]
    assert changed.any() and not (changed & ~nb).any()
    inner = (ids > 0)
    both_fg = np.zeros_like(changed)
    both_fg[:, 1:] |= (ids[:, 1:] != ids[:, :-1]) & inner[:, 1:] & inner[:, :-1]
    g = rng.randn(1, 24, 24, 3)
    dc, dp = R.antialias_bwd(color, rast, pos, tri, g.astype(np.float32), opp)
    for idx in [(0, 5, 7, 1), (0, 12, 12, 0), (0, 20, 3, 2)]:
        fd = _fd(lambda c: R.antialias(c, rast, pos, tri, opp), color, g, idx, 1e-2)
        assert abs(fd - dc[idx]) < 2e-3 * max(1, abs(dc[idx]))
    # position gradient (the silhouette / mask-loss gradient): finite differences with the rast buffer held fixed.
    # The upstream gradient is constant per channel: where a pair's blend weight crosses zero the blend moves from one
    # pixel of the pair to the other (continuous, but a kink), which a per-pixel random gradient would turn into FD noise.
    gc = np.broadcast_to(rng.randn(3), g.shape).copy()
    dc, dp = R.antialias_bwd(color, rast, pos, tri, gc.astype(np.float32), opp)
    assert np.abs(dp).max() > 0
    worst = 0.0
    for v in range(5):
        for c in (0, 1, 3):
            fd = _fd(lambda p: R.antialias(color, rast, p, tri, opp), pos, gc, (0, v, c), 2e-4)
            worst = max(worst, abs(fd - dp[0, v, c]) / max(1.0, np.abs(dp).max()))
    assert worst < 1e-2, worst


def test_vertex_normals_c_matches_numpy_restatement():
    from oracle import geometry_np as gnp
    rng = np.random.RandomState(1)
    v = rng.randn(2, 30, 3).astype(np.float32)
    f = rng.randint(0, 29, size=(50, 3)).astype(np.int32)        # vertex 29 is isolated -> (0,0,1) fallback
    n_c, nsum = R.vertex_normals(v, f)
    assert np.allclose(n_c, gnp.auto_normals(v, f.astype(np.int64)), atol=1e-6)
    assert np.array_equal(n_c[:, 29], np.array([[0, 0, 1], [0, 0, 1]], np.float32))
    g = rng.randn(2, 30, 3)
    dp = R.vertex_normals_bwd(v, f, nsum, g.astype(np.float32))
    for idx in [(0, 3, 0), (1, 10, 2), (0, 20, 1)]:
        fd = _fd(lambda x: R.vertex_normals(x, f)[0], v, g, idx, 1e-3)
        assert abs(fd - dp[idx]) < 2e-2 * max(1, abs(dp[idx]))


def test_coverage_and_visibility_agree_with_an_independent_fp64_evaluation():
    """Independent cross-check of the restated sampling rule on an extracted DMTet mesh: a brute-force numpy fp64
    evaluation (every pixel centre against every projected triangle: signed-area containment, nearest interpolated z/w)
    must give the same covered set and the same visible triangle as the C restatement, at every pixel that is not within
    rounding distance of an edge or of a depth tie (where fp32 and fp64 may legitimately differ)."""
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    syn = importlib.import_module("3danimals_b200.synthetic")
    from oracle import geometry_np as gnp
    v, t = syn.kuhn_tet_grid(20)
    v = v * np.float32(7.0)
    o = gnp.marching_tets(v, syn.sdf_horse(v, 0.0, 0), t, with_uvs=False)
    verts, faces = o["verts"], o["faces"].astype(np.int32)
    mvp, _, _ = syn.cameras(2, seed=5)
    pos = R.xfm_points(verts[None], mvp)
    H = W = 96
    rast = R.rasterize(pos, faces, (H, W))
    fy, fx = np.meshgrid((np.arange(H) + 0.5) / H * 2 - 1, (np.arange(W) + 0.5) / W * 2 - 1, indexing="ij")   # row 0 is clip y = -1
    centre = np.stack([fx, fy], -1).reshape(-1, 1, 1, 2)                       # [P,1,1,2]
    checked = 0
    for b in range(2):
        P = pos[b].astype(np.float64)
        ndc = (P[:, :2] / P[:, 3:4])[faces]                                   # [F,3,2]
        zw = (P[:, 2] / P[:, 3])[faces]                                       # [F,3]
        cover = np.zeros(H * W, bool)
        sure = np.ones(H * W, bool)
        best = np.full(H * W, -1)
        for p0 in range(0, H * W, 1024):                                      # chunks of pixels x all faces
            d = ndc[None] - centre[p0:p0 + 1024]                              # [p,F,3,2]
            e = d[:, :, [1, 2, 0]]
            cr = d[..., 0] * e[..., 1] - d[..., 1] * e[..., 0]               # [p,F,3] signed areas
            area = cr.sum(-1)
            inside = ((cr >= 0).all(-1) | (cr <= 0).all(-1)) & (np.abs(area) > 1e-14)
            edge_close = (np.abs(cr).min(-1) < 1e-6 * np.abs(area).clip(1e-14)) & (np.abs(area) > 1e-14) & \
                         (((cr >= -1e-9).all(-1)) | ((cr <= 1e-9).all(-1)))
            wgt = np.abs(cr[..., [1, 2, 0]]) / np.abs(area)[..., None].clip(1e-300)
            z = np.where(inside, (wgt * zw[None]).sum(-1), np.inf)            # screen-space interpolation of z/w is exact
            zs = np.sort(z, -1)
            cover[p0:p0 + 1024] = inside.any(-1)
            best[p0:p0 + 1024] = np.where(inside.any(-1), z.argmin(-1), -1)
            with np.errstate(invalid="ignore"):
                tie = (zs[:, 1] - zs[:, 0]) < 1e-6
            sure[p0:p0 + 1024] = ~edge_close.any(-1) & ~tie
        ours = rast[b, ..., 3].reshape(-1).astype(np.int64) - 1
        m = sure
        assert np.array_equal(ours[m] >= 0, cover[m])
        assert np.array_equal(ours[m], best[m])
        checked += int(m.sum())
        assert cover[m].sum() > 800
    assert checked > 2 * H * W * 0.9


def test_rasterizer_is_invariant_to_face_order_and_vertex_rotation():
    """Two more properties the restated fill rule must have (nothing pins it against nvdiffrast itself):
    visibility does not depend on the order of the triangle list (away from exact depth ties), and rotating the vertex
    order inside each triangle keeps coverage and depth and rotates the barycentrics accordingly."""
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    syn = importlib.import_module("3danimals_b200.synthetic")
    from oracle import geometry_np as gnp
    v, t = syn.kuhn_tet_grid(16)
    v = v * np.float32(7.0)
    o = gnp.marching_tets(v, syn.sdf_horse(v, 0.01, 3), t, with_uvs=False)
    verts, faces = o["verts"], o["faces"].astype(np.int32)
    mvp, _, _ = syn.cameras(2, seed=9)
    pos = R.xfm_points(verts[None], mvp)
    H = W = 80
    base = R.rasterize(pos, faces, (H, W))
    ids = base[..., 3].astype(np.int64) - 1
    # (1) random permutation of the triangle list
    perm = np.random.RandomState(1).permutation(len(faces))
    r2 = R.rasterize(pos, faces[perm], (H, W))
    ids2 = r2[..., 3].astype(np.int64) - 1
    back = np.where(ids2 >= 0, perm[np.clip(ids2, 0, None)], -1)
    same = back == ids
    assert np.array_equal(ids >= 0, ids2 >= 0)                          # identical coverage
    assert same.mean() > 0.999                                          # identical winner except exact depth ties (id tie-break)
    assert np.allclose(base[..., :3][same], r2[..., :3][same], atol=0)  # and then identical barycentrics / depth bits
    # (2) rotate the vertex order of every triangle: (v0,v1,v2) -> (v1,v2,v0)
    r3 = R.rasterize(pos, faces[:, [1, 2, 0]], (H, W))
    assert np.array_equal(r3[..., 3], base[..., 3])
    cov = ids >= 0
    u, vv = base[..., 0][cov], base[..., 1][cov]
    w = 1.0 - u - vv
    # new weights of (new v0, new v1) = old weights of (v1, v2)
    assert np.abs(r3[..., 0][cov] - vv).max() < 2e-5 and np.abs(r3[..., 1][cov] - w).max() < 2e-5
    assert np.abs(r3[..., 2][cov] - base[..., 2][cov]).max() < 1e-6


def test_antialias_is_linear_in_colour_and_preserves_constants():
    """For fixed geometry the restated antialias is a linear map of the colour image whose rows sum to one: a constant image
    comes back unchanged and aa(a c1 + b c2) = a aa(c1) + b aa(c2)."""
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    syn = importlib.import_module("3danimals_b200.synthetic")
    from oracle import geometry_np as gnp
    v, t = syn.kuhn_tet_grid(12)
    v = v * np.float32(7.0)
    o = gnp.marching_tets(v, syn.sdf_horse(v, 0.0, 0), t, with_uvs=False)
    verts, faces = o["verts"], o["faces"].astype(np.int32)
    mvp, _, _ = syn.cameras(2, seed=4)
    pos = R.xfm_points(verts[None], mvp)
    H = W = 64
    rast = R.rasterize(pos, faces, (H, W))
    opp = R.edge_adjacency(faces, verts.shape[0])
    rng = np.random.RandomState(2)
    c1, c2 = rng.rand(2, H, W, 3).astype(np.float32), rng.rand(2, H, W, 3).astype(np.float32)
    a1, a2 = R.antialias(c1, rast, pos, faces, opp), R.antialias(c2, rast, pos, faces, opp)
    assert np.abs(a1 - c1).max() > 0.02                                   # the scene does blend
    const = np.full((2, H, W, 3), 0.37, np.float32)
    assert np.abs(R.antialias(const, rast, pos, faces, opp) - const).max() < 1e-6
    mix = R.antialias((0.25 * c1 + 1.5 * c2).astype(np.float32), rast, pos, faces, opp)
    assert np.abs(mix - (0.25 * a1 + 1.5 * a2)).max() < 1e-5


def test_barycentrics_are_perspective_correct_fp64():
    """rast (u, v) = perspective-correct weights of triangle vertices 0 and 1: an fp64 evaluation from screen-space areas and
    the clip-space w of each vertex (b_i = (l_i / w_i) / sum_j (l_j / w_j)) agrees with the restatement on every covered pixel,
    and interpolating the clip-space z and w with them reproduces channel 2 (z/w)."""
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    syn = importlib.import_module("3danimals_b200.synthetic")
    from oracle import geometry_np as gnp
    v, t = syn.kuhn_tet_grid(12)
    v = v * np.float32(7.0)
    o = gnp.marching_tets(v, syn.sdf_horse(v, 0.0, 0), t, with_uvs=False)
    verts, faces = o["verts"], o["faces"].astype(np.int32)
    mvp, _, _ = syn.cameras(1, seed=8, fov_deg=60.0, z_offset=4.0)          # strong perspective: w varies by ~2x over the mesh
    pos = R.xfm_points(verts[None], mvp)
    H = W = 128
    rast = R.rasterize(pos, faces, (H, W))[0]
    P = pos[0].astype(np.float64)
    ys, xs = np.nonzero(rast[..., 3] > 0)
    assert len(ys) > 500 and P[:, 3].max() / P[:, 3].min() > 1.3
    f = faces[rast[ys, xs, 3].astype(np.int64) - 1]                          # [n,3]
    c = np.stack([(xs + 0.5) / W * 2 - 1, (ys + 0.5) / H * 2 - 1], -1)       # pixel centres, row 0 is clip y = -1
    p = P[f]                                                                 # [n,3,4]
    ndc = p[..., :2] / p[..., 3:4]
    d = ndc - c[:, None]
    e = d[:, [1, 2, 0]]
    cr = d[..., 0] * e[..., 1] - d[..., 1] * e[..., 0]                        # area opposite vertex (i+2)
    lam = cr[:, [1, 2, 0]] / cr.sum(-1, keepdims=True)                       # screen-space weights of v0, v1, v2
    b = lam / p[..., 3]
    b = b / b.sum(-1, keepdims=True)                                         # perspective-correct weights
    ok = np.abs(cr).min(-1) > 1e-7 * np.abs(cr.sum(-1))                      # away from edges
    assert ok.mean() > 0.95
    assert np.abs(rast[ys, xs, 0] - b[:, 0])[ok].max() < 2e-4 and np.abs(rast[ys, xs, 1] - b[:, 1])[ok].max() < 2e-4
    zw = (b * p[..., 2]).sum(-1) / (b * p[..., 3]).sum(-1)
    assert np.abs(rast[ys, xs, 2] - zw)[ok].max() < 2e-5


def test_antialias_of_axis_aligned_edges_is_the_exact_area_coverage():
    """An analytic consequence of the published antialiasing rule that nothing in the restatement hard-codes: for a straight
    axis-aligned silhouette edge crossing between two pixel centres at fraction d from the inside pixel, the pair blend
    (weight |d - 0.5| on the pixel the edge does NOT cut through the centre side of) turns a 0/1 mask into the exact box-filter
    area of each pixel.  Checked on all four sides of a rectangle (two triangles: the diagonal is an interior edge and must
    stay untouched), away from the corners, at w = 1 and at w = 3 (the rule works on post-divide positions)."""
    H, W = 32, 40
    x0, x1, y0, y1 = 7.3, 29.85, 5.62, 24.41                                   # rectangle in pixel units
    cx = lambda p: p / W * 2 - 1
    cy = lambda p: p / H * 2 - 1
    tri = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    ax = np.clip(np.minimum(np.arange(W) + 1, x1) - np.maximum(np.arange(W), x0), 0, 1)     # exact 1-D coverages
    ay = np.clip(np.minimum(np.arange(H) + 1, y1) - np.maximum(np.arange(H), y0), 0, 1)
    area = ay[:, None] * ax[None, :]
    for w in (1.0, 3.0):
        pos = np.array([[[cx(x0), cy(y0), 0.2, 1], [cx(x1), cy(y0), 0.2, 1], [cx(x1), cy(y1), 0.2, 1], [cx(x0), cy(y1), 0.2, 1]]], np.float32) * np.float32(w)
        rast = R.rasterize(pos, tri, (H, W))
        mask = (rast[..., 3:4] > 0).astype(np.float32)
        inside = mask[0, ..., 0] > 0
        want_inside = (np.abs(np.arange(H) + 0.5 - (y0 + y1) / 2)[:, None] < (y1 - y0) / 2) & (np.abs(np.arange(W) + 0.5 - (x0 + x1) / 2)[None, :] < (x1 - x0) / 2)
        assert np.array_equal(inside, want_inside)
        aa = R.antialias(mask, rast, pos, tri)[0, ..., 0]
        # away from the corners (pixels whose row AND column are both cut by an edge see two blends)
        cut_x = (ax > 0) & (ax < 1)
        cut_y = (ay > 0) & (ay < 1)
        near_x = cut_x | np.roll(cut_x, 1) | np.roll(cut_x, -1)
        near_y = cut_y | np.roll(cut_y, 1) | np.roll(cut_y, -1)
        straight = ~(near_y[:, None] & near_x[None, :])
        assert np.abs(aa - area)[straight].max() < 2e-5, np.abs(aa - area)[straight].max()
        touched = np.abs(aa - mask[0, ..., 0]) > 1e-7
        assert touched.sum() >= 2 * (18 + 22) - 8                               # one pixel of every straddling pair along the four sides
        assert not touched[8:22, 10:27].any()                                   # the interior (incl. the diagonal) is untouched


def test_restatement_regression_pins():
    """tests/golden/raster_scenes.npz freezes the restatement's outputs (ids, barycentrics, interpolation, antialiasing, all
    gradients) on the analytic scenes and one extracted mesh - regression pins of the yardstick, NOT reference outputs (DESIGN.md
    §2).  Forward results are bit-identical (the C file is built with -ffp-contract=off); gradients accumulate in thread order
    under OpenMP, hence a 1e-6 band."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import golden
    from raster_scenes import ANALYTIC, RESOLUTIONS
    g = golden("raster_scenes.npz")
    cases = [(name, res, pos, tri) for name, (pos, tri) in sorted(ANALYTIC.items()) for res in RESOLUTIONS]
    cases.append(("horse_mesh", (64, 64), g["horse_mesh:64x64:pos"], g["horse_mesh:64x64:tri"]))
    covered = 0
    for name, res, pos, tri in cases:
        k = "%s:%dx%d:" % (name, res[0], res[1])
        rast = R.rasterize(pos, tri, res)
        assert np.array_equal(rast, g[k + "rast"]), k
        covered += int((rast[..., 3] > 0).sum())
        col = R.interpolate(g[k + "attr"], rast, tri)
        assert np.array_equal(col, g[k + "col"]), k
        assert np.array_equal(R.antialias(col, rast, pos, tri), g[k + "aa"]), k
        d_attr, d_rast = R.interpolate_bwd(g[k + "attr"], rast, tri, g[k + "g"])
        d_col, d_pos_aa = R.antialias_bwd(col, rast, pos, tri, g[k + "g"])
        d_pos_r = R.rasterize_bwd(pos, tri, rast, g[k + "d_rast"])
        for got, key in ((d_attr, "d_attr"), (d_rast, "d_rast"), (d_col, "d_col"), (d_pos_aa, "d_pos_aa"), (d_pos_r, "d_pos_r")):
            want = g[k + key]
            assert np.abs(got - want).max() <= 1e-6 * max(1.0, np.abs(want).max()), k + key
    assert covered > 4000
