"""ctypes/numpy front-end of oracle/raster_ref.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.isfile(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            path = _build.build()
        _lib = C.CDLL(path)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.c_void_p)


def num_threads():
    return int(lib().orc_num_threads())


def xfm_points(pts, mtx):
    pts, pp = _f(pts); mtx, mp = _f(mtx)
    B, Bp, V = mtx.shape[0], pts.shape[0], pts.shape[1]
    out = np.empty((B, V, 4), np.float32)
    lib().orc_xfm_points_fwd(pp, mp, B, Bp, V, out.ctypes.data_as(C.c_void_p))
    return out


def xfm_points_bwd(pts, mtx, d_out):
    pts, pp = _f(pts); mtx, mp = _f(mtx); d_out, gp = _f(d_out)
    B, Bp, V = mtx.shape[0], pts.shape[0], pts.shape[1]
    d_pts = np.zeros_like(pts); d_mtx = np.zeros_like(mtx)
    lib().orc_xfm_points_bwd(pp, mp, gp, B, Bp, V, d_pts.ctypes.data_as(C.c_void_p), d_mtx.ctypes.data_as(C.c_void_p))
    return d_pts, d_mtx


def rasterize(pos, tri, resolution):
    pos, pp = _f(pos); tri, tp = _i(tri)
    B, V = pos.shape[:2]; F = tri.shape[0]; H, W = resolution
    rast = np.empty((B, H, W, 4), np.float32)
    lib().orc_rasterize_fwd(pp, tp, B, V, F, H, W, rast.ctypes.data_as(C.c_void_p))
    return rast


def rasterize_bwd(pos, tri, rast, d_rast):
    pos, pp = _f(pos); tri, tp = _i(tri); rast, rp = _f(rast); d_rast, gp = _f(d_rast)
    B, V = pos.shape[:2]; F = tri.shape[0]; H, W = rast.shape[1:3]
    d_pos = np.zeros_like(pos)
    lib().orc_rasterize_bwd(pp, tp, rp, gp, B, V, F, H, W, d_pos.ctypes.data_as(C.c_void_p))
    return d_pos


def interpolate(attr, rast, tri):
    attr, ap = _f(attr); rast, rp = _f(rast); tri, tp = _i(tri)
    Ba, V, Cc = attr.shape; B, H, W = rast.shape[:3]; F = tri.shape[0]
    out = np.empty((B, H, W, Cc), np.float32)
    lib().orc_interpolate_fwd(ap, rp, tp, B, Ba, V, F, H, W, Cc, out.ctypes.data_as(C.c_void_p))
    return out


def interpolate_bwd(attr, rast, tri, d_out):
    attr, ap = _f(attr); rast, rp = _f(rast); tri, tp = _i(tri); d_out, gp = _f(d_out)
    Ba, V, Cc = attr.shape; B, H, W = rast.shape[:3]; F = tri.shape[0]
    d_attr = np.zeros_like(attr); d_rast = np.zeros_like(rast)
    lib().orc_interpolate_bwd(ap, rp, tp, gp, B, Ba, V, F, H, W, Cc, d_attr.ctypes.data_as(C.c_void_p),
                              d_rast.ctypes.data_as(C.c_void_p))
    return d_attr, d_rast


def edge_adjacency(tri, V):
    tri, tp = _i(tri)
    opp = np.empty_like(tri)
    lib().orc_edge_adjacency(tp, tri.shape[0], int(V), opp.ctypes.data_as(C.c_void_p))
    return opp


def antialias(color, rast, pos, tri, opp=None):
    color, cp = _f(color); rast, rp = _f(rast); pos, pp = _f(pos); tri, tp = _i(tri)
    B, H, W, Cc = color.shape; V = pos.shape[1]; F = tri.shape[0]
    if opp is None:
        opp = edge_adjacency(tri, V)
    opp, op = _i(opp)
    out = np.empty_like(color)
    lib().orc_antialias_fwd(cp, rp, pp, tp, op, B, V, F, H, W, Cc, out.ctypes.data_as(C.c_void_p))
    return out


def antialias_bwd(color, rast, pos, tri, d_out, opp=None):
    color, cp = _f(color); rast, rp = _f(rast); pos, pp = _f(pos); tri, tp = _i(tri); d_out, gp = _f(d_out)
    B, H, W, Cc = color.shape; V = pos.shape[1]; F = tri.shape[0]
    if opp is None:
        opp = edge_adjacency(tri, V)
    opp, op = _i(opp)
    d_color = np.empty_like(color); d_pos = np.zeros_like(pos)
    lib().orc_antialias_bwd(cp, rp, pp, tp, op, gp, B, V, F, H, W, Cc, d_color.ctypes.data_as(C.c_void_p),
                            d_pos.ctypes.data_as(C.c_void_p))
    return d_color, d_pos


def vertex_normals(v_pos, tri):
    v_pos, vp = _f(v_pos); tri, tp = _i(tri)
    B, V = v_pos.shape[:2]; F = tri.shape[0]
    nsum = np.empty_like(v_pos); nrm = np.empty_like(v_pos)
    lib().orc_vertex_normals_fwd(vp, tp, B, V, F, nsum.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p))
    return nrm, nsum


def vertex_normals_bwd(v_pos, tri, nsum, d_nrm):
    v_pos, vp = _f(v_pos); tri, tp = _i(tri); nsum, sp = _f(nsum); d_nrm, gp = _f(d_nrm)
    B, V = v_pos.shape[:2]; F = tri.shape[0]
    d_pos = np.zeros_like(v_pos)
    lib().orc_vertex_normals_bwd(vp, tp, sp, gp, B, V, F, d_pos.ctypes.data_as(C.c_void_p))
    return d_pos
