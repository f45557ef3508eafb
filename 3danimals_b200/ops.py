"""torch.autograd wrappers over the C-ABI CUDA library (include/b2a.h).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape; every number on the hot path is
produced by a hand-written sm_100a kernel in libb2a.so.  All ops require CUDA fp32 tensors and raise otherwise -
there is no eager / CPU fallback.
"""
import torch

from . import _lib

_i32 = torch.int32


def _L():
    return _lib.lib()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """Raw cudaStream_t of torch's current stream on the current device (kernels are launched where the reference's
    plugin launches: torch_bindings.cpp:170).  The private fast accessor avoids ~20 us of Stream-object churn per call."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


_f32_t = torch.float32


def _f32(t, name):
    try:
        if t.is_cuda and t.dtype is _f32_t and t.is_contiguous():      # the common case: nothing to do
            return t
    except AttributeError:
        pass
    if not torch.is_tensor(t) or not t.is_cuda:
        raise _lib.B2AError("%s must be a CUDA tensor (the B200 hot path has no CPU fallback)" % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _idx32(t, name):
    try:
        if t.is_cuda and t.dtype is _i32 and t.is_contiguous():
            return t
    except AttributeError:
        pass
    if not torch.is_tensor(t) or not t.is_cuda:
        raise _lib.B2AError("%s must be a CUDA tensor" % name)
    if t.dtype != _i32:
        t = t.to(_i32)
    return t.contiguous()


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# kernels launched by each C-ABI entry point (memsets / memcpys are not kernels) - the source of bench.py's gpu_launches
KERNELS_PER_CALL = {
    "b2a_mt_count": 4, "b2a_mt_emit": 2, "b2a_mt_bwd": 1, "b2a_estimate_bones": 4, "b2a_lbs_bone_transforms": 2, "b2a_lbs_fwd": 1, "b2a_lbs_bwd": 1,
    "b2a_lbs_bone_transforms_bwd": 2, "b2a_vertex_normals_fwd": 2, "b2a_vertex_normals_bwd": 2, "b2a_xfm_points_fwd": 1, "b2a_mlp_pack_weights_many": 1, "b2a_rows_gather": 1, "b2a_rows_scatter": 2, "b2a_composite_up_pool_fwd": 1, "b2a_composite_up_pool_bwd": 1, "b2a_articulation_constraints_fwd": 1, "b2a_articulation_constraints_bwd": 1,
    "b2a_xfm_points_bwd": 1, "b2a_rasterize_fwd": 3, "b2a_rasterize_bwd": 1, "b2a_interpolate_fwd": 1, "b2a_interpolate_bwd": 1,
    "b2a_edge_adjacency": 3, "b2a_antialias_prepare": 2, "b2a_antialias_fwd": 1, "b2a_antialias_bwd": 2,
    "b2a_antialias_pair_fwd": 1, "b2a_antialias_pair_bwd": 1, "b2a_shade_directional_fwd": 1, "b2a_shade_directional_bwd": 1,
    "b2a_analytic_field_fwd": 1, "b2a_analytic_field_bwd": 1, "b2a_composite_up_fwd": 1, "b2a_composite_up_bwd": 1, "b2a_gbuffer_fwd": 2, "b2a_gbuffer_bwd": 2,
    "b2a_render_geometry_fwd": 8, "b2a_render_geometry_bwd": 2,
}


class CallStats:
    """Launch counter and optional per-call CUDA-event timer (events are recorded on the stream the kernels are
    launched on).  bench.py enables timing to measure the raster-backward kernels live inside the timed step."""

    def __init__(self):
        self.launches = 0
        self.calls = {}
        self.timing = False
        self.spin_cycles = 0  # device busy-wait enqueued before each timed call (see _call)
        self.events = []      # (name, tag, start_event, end_event)
        self.tag = ""

    def reset(self):
        self.launches = 0
        self.calls = {}
        self.events = []
        cpp = _lib.cpp_nodes()
        if cpp is not None:
            cpp.reset_stats()

    def count(self, name, launches):
        """A C-ABI call made outside _call (the peer-memory all-reduce in parallel.py)."""
        self.launches += launches
        self.calls[name] = self.calls.get(name, 0) + 1

    def total_launches(self):
        """Kernels launched through the Python nodes and through the C++ autograd layer."""
        cpp = _lib.cpp_nodes()
        return self.launches + (cpp.launches() if cpp is not None else 0)

    def all_calls(self):
        cpp = _lib.cpp_nodes()
        out = dict(self.calls)
        if cpp is not None:
            for k, v in cpp.calls().items():
                out[k] = out.get(k, 0) + v
        return out

    def durations_ms(self):
        """{(name, tag): [ms, ...]} - call after torch.cuda.synchronize()."""
        out = {}
        for name, tag, e0, e1 in self.events:
            out.setdefault((name, tag), []).append(e0.elapsed_time(e1))
        return out


stats = CallStats()


def _cpp():
    """The C++ autograd nodes, unless per-call event timing is on (the Python nodes carry the timing hooks)."""
    return None if stats.timing else _lib.cpp_nodes()


_fn_cache = {}
_SYNC_CALLS = __import__("os").environ.get("B2A_SYNC_CALLS", "0") == "1"


def _call(name, args, tag=None, launches=None):
    fn = _fn_cache.get(name)
    if fn is None:
        fast = _lib.fast()
        fn = _fn_cache[name] = getattr(fast, name, None) or getattr(_L(), name)
    tag = stats.tag if tag is None else tag
    if stats.timing:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        if stats.spin_cycles:
            # keep the stream busy while the host enqueues (event, launches, event): otherwise, in a host-bound step,
            # the start event fires on an idle GPU and the host's enqueue latency is billed to the kernel
            torch.cuda._sleep(stats.spin_cycles)
        e0.record()
        rc = fn(*args)
        e1.record()
        stats.events.append((name, tag, e0, e1))
    else:
        rc = fn(*args)
    stats.launches += KERNELS_PER_CALL[name] if launches is None else launches
    stats.calls[name] = stats.calls.get(name, 0) + 1
    _lib.check(rc)
    if _SYNC_CALLS:      # debugging aid (B2A_SYNC_CALLS=1): surface an asynchronous kernel fault at the call that caused it
        try:
            torch.cuda.synchronize()
        except Exception as e:
            raise _lib.B2AError("%s%s faulted: %s" % (name, args, e)) from e


_size_cache = {}


def _size(fn, *args):
    """Workspace-size query (a pure function of its integer arguments): memoised, the hot path asks the same five
    questions every step."""
    key = (fn.__name__,) + args
    v = _size_cache.get(key)
    if v is None:
        import ctypes
        out = ctypes.c_size_t(0)
        _lib.check(fn(*args, ctypes.byref(out)))
        if len(_size_cache) >= 4096:      # in training V and F change every step: keep the memo bounded
            _size_cache.clear()
        v = _size_cache[key] = out.value
    return v


# ---------------------------------------------------------------------------------------------------------------
# Marching tetrahedra (reference model/geometry/dmtet.py:104-155)
# ---------------------------------------------------------------------------------------------------------------
_BASE_EDGES = (0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3)


class TetGrid:
    """Static per-grid tables, built once when a grid is loaded (reference DMTetGeometry.load_tets, dmtet.py:214-226,
    and generate_edges :283-288): int32 tets and the CSR of the unique sorted (min,max) grid edges in lexicographic
    order - the order torch.unique(dim=0) gives the reference's crossing edges, hence its vertex numbering."""

    def __init__(self, tets, num_verts):
        if not tets.is_cuda:
            raise _lib.B2AError("TetGrid needs CUDA tensors")
        L = _L()
        import ctypes
        dev = tets.device
        self.Vg = int(num_verts)
        self.T = int(tets.shape[0])
        st = _stream()
        # unique sorted (min,max) edges -> CSR by the smaller endpoint, built by the library (b2a_mt_build_edges / b2a_mt_emit_edges:
        # device radix sort + unique + scan; the reference's generate_edges is torch.unique(dim=0) over 6 T index pairs)
        src = tets.contiguous()
        if src.dtype not in (torch.int32, torch.int64):
            src = src.long()
        ws = _workspace(_size(L.b2a_mt_tables_workspace_bytes, self.Vg, self.T), dev)
        start = torch.empty(self.Vg + 1, dtype=_i32, device=dev)
        n_edges = torch.zeros(1, dtype=torch.int64).pin_memory()
        _lib.check(L.b2a_mt_build_edges(_p(src), int(src.dtype == torch.int64), self.Vg, self.T, _p(ws), ws.numel(), _p(start), _p(n_edges), st))
        torch.cuda.current_stream().synchronize()
        self.E = int(n_edges.item())
        self.edge_b = torch.empty(self.E, dtype=_i32, device=dev)
        _lib.check(L.b2a_mt_emit_edges(_p(ws), ws.numel(), self.Vg, self.T, self.E, _p(self.edge_b), st))
        self.edge_start = start
        self.tets = src.to(_i32).contiguous()
        tt, tw = ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(L.b2a_mt_tile_shape(ctypes.byref(tt), ctypes.byref(tw)))
        self.tile_words = torch.empty((self.T + tt.value - 1) // tt.value, tw.value, dtype=_i32, device=dev)
        _lib.check(L.b2a_mt_build_tile_words(_p(self.tets), self.T, _p(self.tile_words), st))
        del ws
        self.workspace = _workspace(_size(L.b2a_mt_workspace_bytes, self.Vg, self.E, self.T), dev)
        # output sizes (V, N1, N2, err): written by the count pass straight into PINNED host memory (zero-copy; under UVA the
        # host pointer is the device pointer), so the one readback of the extraction is an event wait, not a D2H memcpy
        self.counts = torch.zeros(4, dtype=_i32).pin_memory()
        self.counts_ready = torch.cuda.Event()

    def all_edges(self):
        """[E,2] int64, identical to the reference's DMTetGeometry.all_edges."""
        a = torch.repeat_interleave(torch.arange(self.Vg, device=self.tets.device),
                                    (self.edge_start[1:] - self.edge_start[:-1]).long())
        return torch.stack([a, self.edge_b.long()], -1)


class _MarchingTets(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, sdf, grid):
        L = _L()
        pos_c = _f32(pos, "pos").reshape(-1, 3)
        sdf_c = _f32(sdf, "sdf").reshape(-1)
        if pos_c.shape[0] != grid.Vg or sdf_c.shape[0] != grid.Vg:
            raise _lib.B2AError("marching_tets: pos/sdf do not match the grid (%d verts)" % grid.Vg)
        st = _stream()
        _call("b2a_mt_count", (_p(sdf_c), _p(grid.tets), _p(grid.edge_start), _p(grid.edge_b), _p(grid.tile_words), grid.Vg, grid.E, grid.T,
                                  _p(grid.workspace), grid.workspace.numel(), _p(grid.counts), st))
        grid.counts_ready.record()
        grid.counts_ready.synchronize()        # the one device->host hand-off of the extraction (output sizes)
        V, N1, N2, err = grid.counts.tolist()
        if err:
            raise _lib.B2AError("marching_tets: a grid vertex has more than 255 crossing edges")
        Fn = N1 + 2 * N2
        dev = pos_c.device
        verts = torch.empty(V, 3, device=dev)
        vert_edge = torch.empty(V, 2, dtype=_i32, device=dev)
        faces = torch.empty(Fn, 3, dtype=torch.int64, device=dev)
        faces32 = torch.empty(Fn, 3, dtype=_i32, device=dev)
        uv_idx = torch.empty(Fn, 3, dtype=torch.int64, device=dev)
        _call("b2a_mt_emit", (_p(pos_c), _p(sdf_c), _p(grid.tets), _p(grid.edge_start), _p(grid.edge_b), grid.Vg, grid.E,
                                 grid.T, _p(grid.workspace), grid.workspace.numel(), V, N1, N2, _p(verts), _p(vert_edge),
                                 _p(faces32), _p(faces), _p(uv_idx), st))
        ctx.save_for_backward(pos_c, sdf_c, vert_edge)
        ctx.shapes = (pos.shape, sdf.shape)
        ctx.mark_non_differentiable(faces, faces32, uv_idx, vert_edge)
        return verts, faces, uv_idx, faces32, vert_edge

    @staticmethod
    def backward(ctx, d_verts, *_):
        pos_c, sdf_c, vert_edge = ctx.saved_tensors
        need_pos = ctx.needs_input_grad[0]
        d_sdf = torch.zeros_like(sdf_c)
        d_pos = torch.zeros_like(pos_c) if need_pos else None
        V = vert_edge.shape[0]
        if V > 0:
            g = _f32(d_verts, "d_verts")
            _call("b2a_mt_bwd", (_p(pos_c), _p(sdf_c), _p(vert_edge), _p(g), V, _p(d_sdf), _p(d_pos), _stream()))
        return (d_pos.reshape(ctx.shapes[0]) if need_pos else None), d_sdf.reshape(ctx.shapes[1]), None


def marching_tets(pos, sdf, grid):
    """-> verts [V,3] f32 (differentiable w.r.t. sdf and pos), faces [F,3] i64, uv_idx [F,3] i64, faces_i32, vert_edge."""
    return _MarchingTets.apply(pos, sdf, grid)


# ---------------------------------------------------------------------------------------------------------------
# Linear blend skinning (reference model/geometry/skinning.py:369-439)
# ---------------------------------------------------------------------------------------------------------------
def chain_tables(kinematic_tree, num_bones, device):
    """kinematic_tree: list of (bone_id, [dependent bone ids]) (skinning.py:25-46).  For bone k the product runs over
    ancestors(k) in list order (root first) followed by k (skinning.py:389-417).  -> chain_ptr [K+1], chain_ids."""
    chains = {}
    for bone_id, _ in kinematic_tree:
        chains[int(bone_id)] = [int(p) for p, children in kinematic_tree if bone_id in children] + [int(bone_id)]
    ptr, ids = [0], []
    for k in range(num_bones):
        c = chains.get(k, [k])
        if len(c) > 16:
            raise _lib.B2AError("kinematic chain deeper than 16 is not supported")
        ids += c
        ptr.append(len(ids))
    return (torch.tensor(ptr, dtype=_i32, device=device), torch.tensor(ids, dtype=_i32, device=device))


class _LBS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v_pos, bones, angles, chain_ptr, chain_ids, temperature, want_weights, bf=None):
        L = _L()
        v_pos = _f32(v_pos, "v_pos"); bones = _f32(bones, "bones"); angles = _f32(angles, "angles")
        # bf = (batch, frames): the caller's [B|1,F|1,...] tensors are flattened HERE (and the outputs / gradients un-flattened
        # below), outside the autograd tape - four recorded view ops fewer per direction than reshaping around the node
        ctx.shapes = (v_pos.shape, angles.shape) if bf is not None else None
        if bf is not None:
            v_pos = v_pos.reshape(-1, v_pos.shape[-2], 3); bones = bones.reshape(-1, bones.shape[-3], 2, 3)
            angles = angles.reshape(-1, angles.shape[-2], 3)
        B, K = angles.shape[0], angles.shape[1]
        Bv, V = v_pos.shape[0], v_pos.shape[1]
        Bb = bones.shape[0]
        dev = v_pos.device
        st = _stream()
        T_local = torch.empty(B, K, 12, device=dev)
        G = torch.empty(B, K, 12, device=dev)
        posed = torch.empty(B, K, 2, 3, device=dev)
        _call("b2a_lbs_bone_transforms", (_p(bones), _p(angles), _p(chain_ptr), _p(chain_ids), B, Bb, K, _p(T_local), _p(G),
                                             _p(posed), st))
        out = torch.empty(B, V, 3, device=dev)
        Bw = max(Bv, Bb)
        weights = torch.empty(K, Bw, V, device=dev) if want_weights else None
        _call("b2a_lbs_fwd", (_p(v_pos), _p(bones), _p(G), B, Bv, Bb, K, V, 1.0 / float(temperature), _p(out), _p(weights), st))
        ctx.save_for_backward(v_pos, bones, angles, chain_ptr, chain_ids, T_local, G)
        ctx.inv_t = 1.0 / float(temperature)
        if weights is None:
            weights = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(weights)
        if bf is not None:
            out, posed = out.view(bf[0], bf[1], V, 3), posed.view(bf[0], bf[1], K, 2, 3)
        return out, posed, weights

    @staticmethod
    def backward(ctx, d_out, d_posed, _):
        L = _L()
        v_pos, bones, angles, chain_ptr, chain_ids, T_local, G = ctx.saved_tensors
        if ctx.shapes is not None:
            d_out = d_out.reshape(-1, d_out.shape[-2], 3) if d_out is not None else None
            d_posed = d_posed.reshape(-1, d_posed.shape[-3], 2, 3) if d_posed is not None else None
        B, K = angles.shape[0], angles.shape[1]
        Bv, V = v_pos.shape[0], v_pos.shape[1]
        Bb = bones.shape[0]
        dev = v_pos.device
        st = _stream()
        # one zero fill for the three accumulators (d_G, d_T, d_v): the step is host-bound, every launch counts
        n, need_v = B * K * 12, ctx.needs_input_grad[0]
        zbuf = torch.zeros(2 * n + (Bv * V * 3 if need_v else 0), device=dev)
        d_G = zbuf[:n].view(B, K, 12)
        d_v = zbuf[2 * n:].view(Bv, V, 3) if need_v else None
        if d_out is not None and V > 0:
            g = _f32(d_out, "d_out")
            _call("b2a_lbs_bwd", (_p(v_pos), _p(bones), _p(G), _p(g), B, Bv, Bb, K, V, ctx.inv_t, _p(d_v), _p(d_G), st))
        d_angles = None
        if ctx.needs_input_grad[2]:
            d_T = zbuf[n:2 * n].view(B, K, 12)
            d_angles = torch.empty(B, K, 3, device=dev)
            gp = _f32(d_posed, "d_posed") if d_posed is not None else None
            _call("b2a_lbs_bone_transforms_bwd", (_p(bones), _p(angles), _p(chain_ptr), _p(chain_ids), _p(T_local), _p(d_G),
                                                     _p(gp), B, Bb, K, _p(d_T), _p(d_angles), st))
        if ctx.shapes is not None:
            d_v = d_v.view(ctx.shapes[0]) if d_v is not None else None
            d_angles = d_angles.view(ctx.shapes[1]) if d_angles is not None else None
        return d_v, None, d_angles, None, None, None, None, None


def lbs_bf(v_pos, bones, angles, chain_ptr, chain_ids, temperature=1.0):
    """The [batch, frames] form skinning() receives: v_pos [B|1,F|1,V,3] (both 1 or both full), bones [B|1,F|1,K,2,3],
    angles [B,F,K,3] -> posed verts [B,F,V,3], posed bones [B,F,K,2,3]."""
    cpp = _cpp()
    if cpp is not None:
        B, Fr, K = angles.shape[0], angles.shape[1], angles.shape[2]
        V = v_pos.shape[-2]
        out, posed = cpp.lbs(v_pos.reshape(-1, V, 3), bones.reshape(-1, K, 2, 3), angles.reshape(-1, K, 3), chain_ptr, chain_ids, float(temperature))
        return out.view(B, Fr, V, 3), posed.view(B, Fr, K, 2, 3)
    out, posed, _ = _LBS.apply(v_pos, bones, angles, chain_ptr, chain_ids, temperature, False, (angles.shape[0], angles.shape[1]))
    return out, posed


def lbs(v_pos, bones, angles, chain_ptr, chain_ids, temperature=1.0, want_weights=False):
    """v_pos [Bv,V,3], bones [Bb,K,2,3], angles [B,K,3] -> posed verts [B,V,3], posed bones [B,K,2,3], weights [K,Bw,V]|None."""
    cpp = _cpp()
    if cpp is not None and not want_weights:
        out, posed = cpp.lbs(v_pos, bones, angles, chain_ptr, chain_ids, float(temperature))
        return out, posed, None
    out, posed, w = _LBS.apply(v_pos, bones, angles, chain_ptr, chain_ids, temperature, want_weights, None)
    return out, posed, (w if want_weights else None)


_eb_ws = {}


def estimate_bones(seq_shape, n_body_bones, n_leg_bones, mode, attach=(-1, -1, -1, -1), want_attach=False, bone_y_threshold=None):
    """seq_shape [B,F,V,3] -> bones [B,F,K,2,3] in one memset + 4 launches, no host sync (reference skinning.py:49-248).
    attach: body-bone index per leg, -1 = auto.  want_attach: also return the device int32[4] with instance 0's choice."""
    L = _L()
    x = _f32(seq_shape, "seq_shape")
    B, Fr, V = x.shape[0], x.shape[1], x.shape[2]
    K = n_body_bones + (4 * n_leg_bones if n_leg_bones > 0 else 0)
    dev = x.device
    ws = _eb_ws.get(dev)
    if ws is None:
        ws = _eb_ws[dev] = _workspace(_size(L.b2a_estimate_bones_workspace_bytes), dev)
    bones = torch.empty(B, Fr, K, 2, 3, device=dev)
    att = torch.empty(4, dtype=_i32, device=dev) if want_attach else None
    qy = 0.0 if bone_y_threshold is None else float(bone_y_threshold)
    _call("b2a_estimate_bones", (_p(x), B * Fr, V, int(n_body_bones), int(n_leg_bones), int(mode), qy, int(attach[0]), int(attach[1]),
                                     int(attach[2]), int(attach[3]), _p(ws), ws.numel(), _p(bones), _p(att), None, _stream()),
          launches=(7 if qy > 0 else 4) if n_leg_bones > 0 else 1)
    return (bones, att) if want_attach else bones


# ---------------------------------------------------------------------------------------------------------------
# Vertex normals (reference model/render/mesh.py:276-304)
# ---------------------------------------------------------------------------------------------------------------
class _VertexNormals(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v_pos, tri):
        v_pos = _f32(v_pos, "v_pos")
        B, V = v_pos.shape[0], v_pos.shape[1]
        F = tri.shape[0]
        nsum = torch.empty(B, V, 4, device=v_pos.device)
        nrm = torch.empty_like(v_pos)
        _call("b2a_vertex_normals_fwd", (_p(v_pos), _p(tri), B, V, F, _p(nsum), _p(nrm), _stream()))
        ctx.save_for_backward(v_pos, tri, nsum)
        return nrm

    @staticmethod
    def backward(ctx, g):
        v_pos, tri, nsum = ctx.saved_tensors
        B, V = v_pos.shape[0], v_pos.shape[1]
        g = _f32(g, "d_nrm")
        scratch = torch.empty_like(nsum)
        d_pos = torch.zeros_like(v_pos)
        _call("b2a_vertex_normals_bwd", (_p(v_pos), _p(tri), _p(nsum), _p(g), B, V, tri.shape[0], _p(scratch), _p(d_pos),
                                               _stream()))
        return d_pos, None


def vertex_normals(v_pos, tri):
    """v_pos [B,V,3], tri [F,3] -> smooth area-weighted vertex normals [B,V,3]."""
    cpp = _cpp()
    if cpp is not None:
        return cpp.vertex_normals(v_pos, _idx32(tri, "tri"))
    return _VertexNormals.apply(v_pos, _idx32(tri, "tri"))


# ---------------------------------------------------------------------------------------------------------------
# Clip transform (reference model/render/renderutils/ops.py:524-525)
# ---------------------------------------------------------------------------------------------------------------
class _XfmPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pts, mtx):
        pts = _f32(pts, "points"); mtx = _f32(mtx, "matrix")
        B, Bp, V = mtx.shape[0], pts.shape[0], pts.shape[1]
        out = torch.empty(B, V, 4, device=pts.device)
        _call("b2a_xfm_points_fwd", (_p(pts), _p(mtx), B, Bp, V, _p(out), _stream()))
        ctx.save_for_backward(pts, mtx)
        return out

    @staticmethod
    def backward(ctx, g):
        pts, mtx = ctx.saved_tensors
        B, Bp, V = mtx.shape[0], pts.shape[0], pts.shape[1]
        g = _f32(g, "d_out")
        d_pts = torch.zeros_like(pts) if ctx.needs_input_grad[0] else None
        d_mtx = torch.zeros_like(mtx) if ctx.needs_input_grad[1] else None
        _call("b2a_xfm_points_bwd", (_p(pts), _p(mtx), _p(g), None, 0, B, Bp, V, _p(d_pts), _p(d_mtx), _stream()))
        return d_pts, d_mtx


def xfm_points(points, matrix):
    """points [B or 1,V,3], matrix [B,4,4] -> homogeneous clip-space points [B,V,4]."""
    return _XfmPoints.apply(points, matrix)


# ---------------------------------------------------------------------------------------------------------------
# Rasterize / interpolate / antialias (nvdiffrast.torch ops used by reference model/render/render.py:24,264,292,351)
# ---------------------------------------------------------------------------------------------------------------
class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, tri, H, W, want_cov):
        L = _L()
        pos = _f32(pos, "pos")
        B, V = pos.shape[0], pos.shape[1]
        F = tri.shape[0]
        ws = _workspace(_size(L.b2a_rasterize_workspace_bytes, B, F, H, W), pos.device)
        rast = torch.empty(B, H, W, 4, device=pos.device)
        cov_list = torch.empty((B * H * W if want_cov else 0, 4), dtype=_i32, device=pos.device)
        cov_count = torch.empty(1 if want_cov else 0, dtype=_i32, device=pos.device)
        _call("b2a_rasterize_fwd", (_p(pos), _p(tri), B, V, F, H, W, _p(ws), ws.numel(), _p(rast), _p(cov_list) if want_cov else None,
                                        _p(cov_count) if want_cov else None, _stream()))
        ctx.save_for_backward(pos, tri, rast)
        ctx.mark_non_differentiable(cov_list, cov_count)
        return rast, cov_list, cov_count

    @staticmethod
    def backward(ctx, g, *_):
        pos, tri, rast = ctx.saved_tensors
        B, V = pos.shape[0], pos.shape[1]
        H, W = rast.shape[1], rast.shape[2]
        g = _f32(g, "d_rast")
        d_pos = torch.zeros_like(pos)
        _call("b2a_rasterize_bwd", (_p(pos), _p(tri), _p(rast), _p(g), B, V, tri.shape[0], H, W, _p(d_pos), _stream()))
        return d_pos, None, None, None, None


def rasterize(pos, tri, resolution, with_coverage=False):
    """pos [B,V,4] clip space, tri [F,3], resolution (H,W) -> rast [B,H,W,4] = (u, v, z/w, triangle_id+1).
    with_coverage: also return (cov_list int32 [B*H*W,4], cov_count int32 [1]) - the compact list of covered pixels
    (flat pixel index, vertex ids of the visible triangle) that lets ops.gbuffer's backward run dense warps."""
    if pos.dim() != 3 or pos.shape[-1] != 4:
        raise _lib.B2AError("rasterize: pos must be [B,V,4] (instanced mode)")
    rast, cl, cc = _Rasterize.apply(pos, _idx32(tri, "tri"), int(resolution[0]), int(resolution[1]), bool(with_coverage))
    return (rast, (cl, cc)) if with_coverage else rast


class _Interpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tri):
        attr = _f32(attr, "attr"); rast = _f32(rast, "rast")
        B, H, W = rast.shape[0], rast.shape[1], rast.shape[2]
        Ba, V, Cc = attr.shape
        out = torch.empty(B, H, W, Cc, device=attr.device)
        _call("b2a_interpolate_fwd", (_p(attr), _p(rast), _p(tri), B, Ba, V, tri.shape[0], H, W, Cc, _p(out), _stream()))
        ctx.save_for_backward(attr, rast, tri)
        return out

    @staticmethod
    def backward(ctx, g):
        attr, rast, tri = ctx.saved_tensors
        B, H, W = rast.shape[0], rast.shape[1], rast.shape[2]
        Ba, V, Cc = attr.shape
        g = _f32(g, "d_out")
        d_attr = torch.zeros_like(attr) if ctx.needs_input_grad[0] else None
        d_rast = torch.empty_like(rast) if ctx.needs_input_grad[1] else None
        _call("b2a_interpolate_bwd", (_p(attr), _p(rast), _p(tri), _p(g), B, Ba, V, tri.shape[0], H, W, Cc, _p(d_attr),
                                            _p(d_rast), _stream()))
        return d_attr, d_rast, None


def interpolate(attr, rast, tri):
    """attr [B or 1,V,C], rast [B,H,W,4], tri [F,3] -> [B,H,W,C]."""
    if attr.dim() != 3 or (attr.shape[0] != 1 and attr.shape[0] != rast.shape[0]):
        raise _lib.B2AError("interpolate: attr must be [B or 1,V,C]")
    return _Interpolate.apply(attr, rast, _idx32(tri, "tri"))


def edge_adjacency(tri, num_verts):
    """tri [F,3] -> opp [F,3] int32: the vertex opposite each edge in the neighbouring triangle (-1: boundary)."""
    L = _L()
    tri = _idx32(tri, "tri")
    F = tri.shape[0]
    ws = _workspace(_size(L.b2a_edge_adjacency_workspace_bytes, F), tri.device)
    opp = torch.empty(F, 3, dtype=_i32, device=tri.device)
    _call("b2a_edge_adjacency", (_p(tri), F, int(num_verts), _p(ws), ws.numel(), _p(opp), _stream()))
    return opp


class _Antialias(torch.autograd.Function):
    """composite=False: plain antialias(color[B,H,W,C]).  composite=True: color is [B,H,W,C-1] and the blended input is
    lerp(bg, [color,1], id>0) (render.py:258-262), never materialised.  `keep` = how many leading output channels the
    caller uses (the rest are sliced off, render.py:320-331), so their gradient is known to be zero."""

    @staticmethod
    def forward(ctx, color, bg, rast, pos, tri, opp, composite, keep, aa_ctx=None):
        color = _f32(color, "color"); rast = _f32(rast, "rast"); pos = _f32(pos, "pos")
        bg = _f32(bg, "background") if bg is not None else None
        B, H, W = rast.shape[0], rast.shape[1], rast.shape[2]
        Cc = color.shape[-1] + (1 if composite else 0)
        if color.shape[:3] != rast.shape[:3]:
            raise _lib.B2AError("antialias: color %s does not match rast %s" % (tuple(color.shape), tuple(rast.shape)))
        Bg = 1
        if bg is not None:
            if bg.shape[1:] != (H, W, Cc) or bg.shape[0] not in (1, B):
                raise _lib.B2AError("antialias: background shape %s, expected [1|B,%d,%d,%d]" % (tuple(bg.shape), H, W, Cc))
            Bg = bg.shape[0]
        out = torch.empty(B, H, W, Cc, device=color.device)
        _call("b2a_antialias_fwd", (_p(color), _p(bg), Bg, int(composite), _p(rast), _p(pos), _p(tri), _p(opp), B, pos.shape[1],
                                          tri.shape[0], H, W, Cc, _p(out), _p(aa_ctx), 0 if aa_ctx is None else aa_ctx.numel(), _stream()),
              tag="C%d" % Cc)
        ctx.save_for_backward(color, bg, rast, pos, tri, opp, aa_ctx)
        ctx.cfg = (bool(composite), Bg, Cc, int(keep))
        if keep < Cc:
            out = out[..., :keep]
        return out

    @staticmethod
    def backward(ctx, g):
        color, bg, rast, pos, tri, opp, aa_ctx = ctx.saved_tensors
        composite, Bg, Cc, keep = ctx.cfg
        B, H, W = rast.shape[0], rast.shape[1], rast.shape[2]
        if g.dtype != torch.float32:
            g = g.float()
        d_color = torch.empty_like(color) if ctx.needs_input_grad[0] else None
        d_pos = torch.zeros_like(pos) if ctx.needs_input_grad[3] else None
        sb, sy, sx, sc = g.stride()
        _call("b2a_antialias_bwd", (_p(color), _p(bg), Bg, int(composite), _p(rast), _p(pos), _p(tri), _p(opp), _p(g), sb, sy, sx,
                                          sc, keep, B, pos.shape[1], tri.shape[0], H, W, Cc, _p(d_color), _p(d_pos), _p(aa_ctx),
                                          0 if aa_ctx is None else aa_ctx.numel(), _stream()), tag="C%d" % Cc,
              launches=1)
        return d_color, None, None, d_pos, None, None, None, None, None


def antialias_prepare(rast, pos, tri, opp):
    """One pass over rast [B,H,W,4] + one pair-analysis launch -> opaque per-render context (coverage / silhouette
    bitmasks, silhouette pixel list, per-pair blend records) shared by all composite_antialias launches of that render.
    Returns None when the shape does not qualify (H*W % 32 != 0)."""
    L = _L()
    rast = _f32(rast, "rast"); pos = _f32(pos, "pos")
    tri = _idx32(tri, "tri"); opp = _idx32(opp, "opp")
    B, H, W = rast.shape[0], rast.shape[1], rast.shape[2]
    if (H * W) % 32 != 0 or tri.shape[0] >= (1 << 28):
        return None
    ws = _workspace(_size(L.b2a_antialias_workspace_bytes, B, H, W), rast.device)
    _call("b2a_antialias_prepare", (_p(rast), _p(pos), _p(tri), _p(opp), B, pos.shape[1], tri.shape[0], H, W, _p(ws), ws.numel(),
                                        _stream()))
    return ws


def antialias(color, rast, pos, tri, opp=None):
    """nvdiffrast.torch.antialias semantics: color [B,H,W,C], rast [B,H,W,4], pos [B,V,4], tri [F,3]."""
    tri = _idx32(tri, "tri")
    if opp is None:
        opp = edge_adjacency(tri, pos.shape[1])
    return _Antialias.apply(color, None, rast, pos, tri, opp, False, color.shape[-1])


def composite_antialias(color, background, rast, pos, tri, opp, antialias_edges=True, keep=None, aa_ctx=None):
    """Fused lerp(background, [color,1], id>0) (+ antialias).  color [B,H,W,C-1]; background [1|B,H,W,C] or None (zeros).
    Returns [B,H,W,keep] (keep defaults to C)."""
    tri = _idx32(tri, "tri")
    Cc = color.shape[-1] + 1
    keep = Cc if keep is None else keep
    if not antialias_edges:
        # composite only: same kernel with an adjacency table that marks every edge interior is not available, so the
        # un-antialiased keys (kd, normal, geo_normal) take the plain torch lerp below; they are logging-only modes.
        alpha = (rast[..., -1:] > 0).float()
        bgt = background if background is not None else torch.zeros(1, *color.shape[1:3], Cc, device=color.device)
        acc = torch.lerp(bgt.expand(color.shape[0], -1, -1, -1), torch.cat((color, torch.ones_like(color[..., :1])), -1), alpha)
        return acc[..., :keep]
    return _Antialias.apply(color, background, rast, pos, tri, opp, True, keep, aa_ctx)


class _AntialiasPair(torch.autograd.Function):
    """composite + antialias of the training pair of keys - a wide one (dino_pred, 16+1 channels) and a narrow one
    (shaded, 3+1) - in ONE launch per direction over the render's prepared context (b2a_antialias_pair_fwd/bwd).  Same
    device code as two _Antialias nodes (bit-identical images); what it removes is a launch ramp + tail per direction,
    one autograd node, one zero fill and the d_pos sum of the two keys."""

    @staticmethod
    def forward(ctx, color_w, color_n, bg_w, bg_n, pos, keep_w, keep_n, aa_ctx, rast, tri, opp, nchw=False):
        color_w = _f32(color_w, "color"); color_n = _f32(color_n, "color"); pos = _f32(pos, "pos")
        bg_w = _f32(bg_w, "background") if bg_w is not None else None
        bg_n = _f32(bg_n, "background") if bg_n is not None else None
        B, H, W = color_w.shape[0], color_w.shape[1], color_w.shape[2]
        Cw, Cn = color_w.shape[-1] + 1, color_n.shape[-1] + 1
        if color_n.shape[:3] != color_w.shape[:3]:
            raise _lib.B2AError("antialias pair: the two keys must share [B,H,W]")
        for bg, Cc in ((bg_w, Cw), (bg_n, Cn)):
            if bg is not None and (bg.shape[1:] != (H, W, Cc) or bg.shape[0] not in (1, B)):
                raise _lib.B2AError("antialias: background shape %s, expected [1|B,%d,%d,%d]" % (tuple(bg.shape), H, W, Cc))
        Bgw = 1 if bg_w is None else bg_w.shape[0]
        Bgn = 1 if bg_n is None else bg_n.shape[0]
        out_w = torch.empty(B, H, W, Cw, device=color_w.device)
        out_n = torch.empty(B, H, W, Cn, device=color_w.device)
        _call("b2a_antialias_pair_fwd", (_p(color_w), _p(bg_w), Bgw, Cw, _p(out_w), _p(color_n), _p(bg_n), Bgn, Cn, _p(out_n), B, H, W,
                                          _p(aa_ctx), aa_ctx.numel(), _stream()), tag="C%d+C%d" % (Cw, Cn))
        ctx.save_for_backward(color_w, color_n, bg_w, bg_n, pos, aa_ctx, rast, tri, opp)
        ctx.cfg = (Bgw, Bgn, Cw, Cn, int(keep_w), int(keep_n), bool(nchw))
        ow = out_w[..., :keep_w] if keep_w < Cw else out_w
        on = out_n[..., :keep_n] if keep_n < Cn else out_n
        if nchw:    # render.py:334 hands NCHW views of the NHWC storage to the caller: made here, outside the autograd tape
            ow, on = ow.permute(0, 3, 1, 2), on.permute(0, 3, 1, 2)
        return ow, on

    @staticmethod
    def backward(ctx, g_w, g_n):
        color_w, color_n, bg_w, bg_n, pos, aa_ctx, rast, tri, opp = ctx.saved_tensors
        Bgw, Bgn, Cw, Cn, keep_w, keep_n, nchw_out = ctx.cfg
        B, H, W = color_w.shape[0], color_w.shape[1], color_w.shape[2]
        g_w = g_w if g_w.dtype == torch.float32 else g_w.float()
        g_n = g_n if g_n.dtype == torch.float32 else g_n.float()
        if nchw_out:    # gradients arrive [B,C,H,W]-shaped: view them [B,H,W,C]-shaped (strides carry the layout)
            g_w, g_n = g_w.permute(0, 2, 3, 1), g_n.permute(0, 2, 3, 1)
        d_color_w = torch.empty_like(color_w)
        d_color_n = torch.empty_like(color_n)
        d_pos = torch.zeros_like(pos) if ctx.needs_input_grad[4] else None
        wsb, wsy, wsx, wsc = g_w.stride()
        nsb, nsy, nsx, nsc = g_n.stride()
        V = pos.shape[1]
        st = _stream()
        nhwc = wsc == 1 and wsx == keep_w and wsy == W * keep_w and wsb % 4 == 0 and g_w.data_ptr() % 16 == 0
        nchw = wsx == 1 and W % 32 == 0
        if (nhwc or nchw) and keep_w == Cw - 1 and keep_n in (Cn, Cn - 1):
            _call("b2a_antialias_pair_bwd", (_p(color_w), _p(bg_w), Bgw, Cw, _p(g_w), wsb, wsy, wsx, wsc, keep_w, _p(d_color_w), _p(color_n),
                                              _p(bg_n), Bgn, Cn, _p(g_n), nsb, nsy, nsx, nsc, keep_n, _p(d_color_n), B, V, H, W, _p(d_pos),
                                              _p(aa_ctx), aa_ctx.numel(), st), tag="C%d+C%d" % (Cw, Cn))
        else:   # gradient layout without a fused instantiation: the two single-key launches, same d_pos accumulator
            for color, bg, Bg, Cc, g, (sb, sy, sx, sc), keep, d_color in (
                    (color_w, bg_w, Bgw, Cw, g_w, (wsb, wsy, wsx, wsc), keep_w, d_color_w),
                    (color_n, bg_n, Bgn, Cn, g_n, (nsb, nsy, nsx, nsc), keep_n, d_color_n)):
                _call("b2a_antialias_bwd", (_p(color), _p(bg), Bg, 1, _p(rast), _p(pos), _p(tri), _p(opp), _p(g), sb, sy, sx, sc, keep, B, V,
                                              tri.shape[0], H, W, Cc, _p(d_color), _p(d_pos), _p(aa_ctx), aa_ctx.numel(), st), tag="C%d" % Cc,
                      launches=1)
        return d_color_w, d_color_n, None, None, d_pos, None, None, None, None, None, None, None


def composite_antialias_pair(color_w, bg_w, keep_w, color_n, bg_n, keep_n, rast, pos, tri, opp, aa_ctx, nchw=False):
    """Fused composite + antialias of a wide key (color_w [B,H,W,16]) and a narrow key (color_n [B,H,W,3]) of one render.
    -> ([B,H,W,keep_w], [B,H,W,keep_n]), or with nchw=True the [B,keep,H,W] views of the same storage that render_mesh returns
    (render.py:334).  Requires the render's prepared context (ops.antialias_prepare)."""
    if not pair_supported(color_w, color_n, aa_ctx):
        raise _lib.B2AError("composite_antialias_pair: unsupported key combination (use composite_antialias per key)")
    cpp = _cpp()
    if (cpp is not None and nchw and color_w.shape[2] % 32 == 0 and int(keep_w) == color_w.shape[-1]
            and int(keep_n) in (color_n.shape[-1], color_n.shape[-1] + 1)):
        ow, on = cpp.antialias_pair(color_w, color_n, bg_w, bg_n, pos, int(keep_w), int(keep_n), aa_ctx)
        return ow, on
    return _AntialiasPair.apply(color_w, color_n, bg_w, bg_n, pos, int(keep_w), int(keep_n), aa_ctx, _f32(rast, "rast"), _idx32(tri, "tri"),
                                _idx32(opp, "opp"), bool(nchw))


def pair_supported(color_w, color_n, aa_ctx):
    return (aa_ctx is not None and color_w.shape[-1] == 16 and color_n.shape[-1] == 3 and color_w.shape[:3] == color_n.shape[:3]
            and color_w.shape[0] * color_w.shape[1] * color_w.shape[2] * 17 < (1 << 31))


class _CompositeUp(torch.autograd.Function):
    """Composite (+ antialias) of a narrow key whose colour lives at the g-buffer resolution [B,H/up,W/up,C-1]: the nearest
    up-sampling of render.py:217-219 happens inside the kernel, the backward writes d_color at low resolution
    (b2a_composite_up_fwd/bwd).  antialias=False: composite only (kd / normal / geo_normal)."""

    @staticmethod
    def forward(ctx, color, bg, pos, up, antialias, keep, aa_ctx, H, W, pool):
        color = _f32(color, "color"); pos = _f32(pos, "pos")
        bg = _f32(bg, "background") if bg is not None else None
        B = color.shape[0]
        Cc = color.shape[-1] + 1
        if color.shape[1] * up != H or color.shape[2] * up != W:
            raise _lib.B2AError("composite_up: colour %s x%d does not match the raster resolution (%d,%d)" % (tuple(color.shape), up, H, W))
        Bg = 1
        if bg is not None:
            if bg.shape[1:] != (H, W, Cc) or bg.shape[0] not in (1, B):
                raise _lib.B2AError("antialias: background shape %s, expected [1|B,%d,%d,%d]" % (tuple(bg.shape), H, W, Cc))
            Bg = bg.shape[0]
        ctx.save_for_backward(color, bg, pos, aa_ctx)
        ctx.cfg = (up, bool(antialias), Bg, Cc, int(keep), H, W, bool(pool))
        if pool:      # the up x up average of the composited image, in the same kernel (no [B,H,W,C] image)
            out = torch.empty(B, H // up, W // up, keep, device=color.device)
            _call("b2a_composite_up_pool_fwd", (_p(color), up, _p(bg), Bg, int(antialias), B, H, W, Cc, int(keep), _p(out), _p(aa_ctx), aa_ctx.numel(),
                                                _stream()), tag="C%d" % Cc)
            return out
        out = torch.empty(B, H, W, Cc, device=color.device)
        _call("b2a_composite_up_fwd", (_p(color), up, _p(bg), Bg, int(antialias), B, H, W, Cc, _p(out), _p(aa_ctx), aa_ctx.numel(), _stream()),
              tag="C%d" % Cc)
        return out[..., :keep] if keep < Cc else out

    @staticmethod
    def backward(ctx, g):
        color, bg, pos, aa_ctx = ctx.saved_tensors
        up, antialias, Bg, Cc, keep, H, W, pool = ctx.cfg
        if g.dtype != torch.float32:
            g = g.float()
        d_color = torch.empty_like(color)
        d_pos = torch.zeros_like(pos) if (ctx.needs_input_grad[2] and antialias) else None
        sb, sy, sx, sc = g.stride()
        _call("b2a_composite_up_pool_bwd" if pool else "b2a_composite_up_bwd",
              (_p(color), up, _p(bg), Bg, int(antialias), _p(g), sb, sy, sx, sc, keep, color.shape[0], pos.shape[1], H, W,
               Cc, _p(d_color), _p(d_pos), _p(aa_ctx), aa_ctx.numel(), _stream()), tag="C%d" % Cc)
        return d_color, None, d_pos, None, None, None, None, None, None, None


class _ScatterRows(torch.autograd.Function):
    """dense[idx[r]] = rows[r], zero elsewhere (b2a_rows_scatter); backward: the gather of the same rows (b2a_rows_gather)."""

    @staticmethod
    def forward(ctx, rows, idx, n):
        rows = _f32(rows, "rows")
        N, C = rows.shape
        out = torch.empty(n, C, device=rows.device)
        _call("b2a_rows_scatter", (_p(rows), _p(idx), N, C, _p(out), n, 1, _stream()))
        ctx.save_for_backward(idx)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = _f32(g, "d_dense")
        out = torch.empty(idx.shape[0], g.shape[1], device=g.device)
        _call("b2a_rows_gather", (_p(g), _p(idx), idx.shape[0], g.shape[1], _p(out), _stream()))
        return out, None, None


def scatter_rows(rows, idx, n):
    """rows [N,C] (the field networks' outputs on the covered pixels), idx [N] int64 unique -> dense [n,C], zero at the other rows."""
    if idx.dtype != torch.int64 or not idx.is_contiguous():
        idx = idx.long().contiguous()
    return _ScatterRows.apply(rows, idx, int(n))


def composite_up_supported(color, aa_ctx):
    return aa_ctx is not None and 1 <= color.shape[-1] <= 3 and color.dtype == torch.float32


def composite_up(color, background, pos, resolution, up=1, antialias_edges=True, keep=None, aa_ctx=None, pool=False):
    """color [B,H/up,W/up,C-1] (C <= 4) -> composited (+ antialiased) [B,H,W,keep] at the raster resolution (H,W); pool: the
    up x up average of that image, [B,H/up,W/up,keep] (the msaa resolve of render.py:322-323, fused)."""
    if not composite_up_supported(color, aa_ctx):
        raise _lib.B2AError("composite_up: needs a prepared context and at most 3 colour channels")
    Cc = color.shape[-1] + 1
    return _CompositeUp.apply(color, background, pos, int(up), bool(antialias_edges), Cc if keep is None else int(keep), aa_ctx,
                              int(resolution[0]), int(resolution[1]), bool(pool) and up > 1)


# ---------------------------------------------------------------------------------------------------------------
# Fused g-buffer (reference model/render/render.py:160-209 + :72-75)
# ---------------------------------------------------------------------------------------------------------------
GB_KEYS = ("pos", "geo_nrm", "shading_nrm", "cam_nrm", "tex_pos")


_gb_acc = {}


def _gb_accumulator(nbytes, device):
    """Per-(device, stream) [B,V,12] vertex-gradient accumulator kept ZEROED between calls (the backward's finalize pass
    re-zeroes what it read), so the steady state issues no memset.  Grown (and zeroed once) on demand."""
    key = (device, _stream())
    ws = _gb_acc.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = _gb_acc[key] = torch.zeros(max(int(nbytes), 16), dtype=torch.uint8, device=device)
    return ws


class _GBuffer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rast, pos_clip, tri, v_pos, v_nrm, prior_pos, w2c, campos, spp, two_sided, want, cov_list, cov_count):
        rast = _f32(rast, "rast"); pos_clip = _f32(pos_clip, "pos_clip"); v_pos = _f32(v_pos, "v_pos"); v_nrm = _f32(v_nrm, "v_nrm")
        prior_pos = _f32(prior_pos, "prior_pos"); w2c = _f32(w2c, "w2c"); campos = _f32(campos, "campos")
        B, V = v_pos.shape[0], v_pos.shape[1]
        H, W = rast.shape[1] // spp, rast.shape[2] // spp
        if rast.shape[0] != B or w2c.shape != (B, 4, 4) or campos.shape != (B, 3) or v_nrm.shape != v_pos.shape:
            raise _lib.B2AError("gbuffer: inconsistent batch shapes")
        outs = [torch.empty(B, H, W, 3, device=rast.device) if k in want else None for k in GB_KEYS]
        packed = _workspace(_size(_L().b2a_gbuffer_pack_bytes, B, prior_pos.shape[0], V), rast.device)
        _call("b2a_gbuffer_fwd", (_p(rast), spp, _p(tri), _p(v_pos), _p(v_nrm), _p(prior_pos), prior_pos.shape[0], _p(w2c),
                                        _p(campos), int(two_sided), B, V, tri.shape[0], H, W, _p(packed), packed.numel(),
                                        *[_p(o) for o in outs], _stream()))
        ctx.save_for_backward(rast, pos_clip, tri, v_pos, v_nrm, prior_pos, w2c, campos, cov_list, cov_count, packed)
        ctx.cfg = (spp, int(two_sided), H, W, tuple(o is not None for o in outs))
        return tuple(o if o is not None else torch.empty(0, device=rast.device) for o in outs)

    @staticmethod
    def backward(ctx, *grads):
        L = _L()
        rast, pos_clip, tri, v_pos, v_nrm, prior_pos, w2c, campos, cov_list, cov_count, packed = ctx.saved_tensors
        spp, two_sided, H, W, present = ctx.cfg
        B, V = v_pos.shape[0], v_pos.shape[1]
        gs = [(_f32(g, "d_gb") if (g is not None and p) else None) for g, p in zip(grads, present)]
        need = ctx.needs_input_grad
        # vertex gradients are WRITTEN by the kernel (from its [B,V,12] accumulator): no zero fills
        d_clip = torch.empty_like(pos_clip) if need[1] else None
        d_v_pos = torch.empty_like(v_pos) if need[3] else None
        d_v_nrm = torch.empty_like(v_nrm) if need[4] else None
        d_prior = torch.empty_like(prior_pos) if need[5] else None
        d_w2c = torch.zeros_like(w2c) if need[6] else None
        d_campos = torch.zeros_like(campos) if need[7] else None
        if any(g is not None for g in gs):
            ws = _gb_accumulator(_size(L.b2a_gbuffer_bwd_workspace_bytes, B, V), rast.device)
            _call("b2a_gbuffer_bwd", (_p(rast), spp, _p(pos_clip), _p(tri), _p(v_pos), _p(v_nrm), _p(prior_pos), prior_pos.shape[0],
                                            _p(w2c), _p(campos), two_sided, B, V, tri.shape[0], H, W, _p(packed), packed.numel(), _p(cov_list),
                                            _p(cov_count), *[_p(g) for g in gs], _p(ws), ws.numel(), 1, _p(d_v_pos), _p(d_v_nrm), _p(d_prior), _p(d_clip),
                                            _p(d_w2c), _p(d_campos), _stream()))
        else:
            for t in (d_clip, d_v_pos, d_v_nrm, d_prior):
                if t is not None:
                    t.zero_()
        return None, d_clip, None, d_v_pos, d_v_nrm, d_prior, d_w2c, d_campos, None, None, None, None, None


def gbuffer(rast, pos_clip, tri, v_pos, v_nrm, prior_pos, w2c, campos, spp=1, two_sided=True, want=("cam_nrm", "tex_pos"), coverage=None):
    """Fused g-buffer pass.  rast [B,H*spp,W*spp,4] (not differentiated: visibility is piecewise constant; barycentric
    gradients go straight to pos_clip).  coverage: the (cov_list, cov_count) pair of ops.rasterize(with_coverage=True)
    (used when spp == 1).  Returns a dict with the requested keys of GB_KEYS, each [B,H,W,3]."""
    cl, cc = coverage if (coverage is not None and int(spp) == 1) else (None, None)
    outs = _GBuffer.apply(rast.detach(), pos_clip, _idx32(tri, "tri"), v_pos, v_nrm, prior_pos, w2c, campos, int(spp), bool(two_sided),
                          tuple(want), cl, cc)
    return {k: o for k, o in zip(GB_KEYS, outs) if k in want}


# ---------------------------------------------------------------------------------------------------------------
# Directional-light shading (reference model/render/light.py:186-193)
# ---------------------------------------------------------------------------------------------------------------
def _rows3(t):
    """-> (tensor to keep alive, row stride in floats) for a [..., 3] fp32 CUDA tensor whose last dim is dense and whose
    leading dims are uniformly strided (a channel slice of an NHWC tensor qualifies); anything else is made contiguous."""
    if t.dtype == torch.float32 and t.stride(-1) == 1 and t.dim() >= 2:
        rs = t.stride(-2)
        ok = rs >= 3
        n = t.shape[-2]
        for d in range(t.dim() - 3, -1, -1):
            ok = ok and (t.shape[d] == 1 or t.stride(d) == rs * n)
            n *= t.shape[d]
        if ok:
            return t, rs
    return t.float().contiguous(), 3


class _ShadeDirectional(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kd, normal, light):
        if not (kd.is_cuda and normal.is_cuda and light.is_cuda):
            raise _lib.B2AError("shade_directional needs CUDA tensors (the B200 hot path has no CPU fallback)")
        normal = _f32(normal, "normal"); light = _f32(light, "light_params")
        kd_t, ks = _rows3(kd)
        B = normal.shape[0]
        HW = normal.numel() // (3 * B)
        Bl = light.shape[0]
        if kd.shape != normal.shape or light.shape[-1] != 5 or Bl not in (1, B):
            raise _lib.B2AError("shade_directional: kd %s, normal %s, light %s" % (tuple(kd.shape), tuple(normal.shape), tuple(light.shape)))
        shaded = torch.empty_like(normal)
        shading = torch.empty(*normal.shape[:-1], 1, device=normal.device)
        _call("b2a_shade_directional_fwd", (_p(kd_t), ks, _p(normal), _p(light), Bl, B, HW, _p(shaded), _p(shading), _stream()))
        ctx.save_for_backward(kd_t, normal, light)
        ctx.ks = ks
        ctx.kd_shape = kd.shape
        return shaded, shading

    @staticmethod
    def backward(ctx, g_shaded, g_shading):
        kd_t, normal, light = ctx.saved_tensors
        B = normal.shape[0]
        HW = normal.numel() // (3 * B)
        need = ctx.needs_input_grad
        if g_shaded is None:
            g_shaded = torch.zeros_like(normal)
        g_shaded = _f32(g_shaded, "d_shaded")
        g_shading = _f32(g_shading, "d_shading") if g_shading is not None else None
        d_kd = torch.empty(ctx.kd_shape, device=normal.device) if need[0] else None
        d_n = torch.empty_like(normal) if need[1] else None
        d_l = torch.zeros_like(light) if need[2] else None
        _call("b2a_shade_directional_bwd", (_p(kd_t), ctx.ks, _p(normal), _p(light), light.shape[0], _p(g_shaded), _p(g_shading), B, HW, _p(d_kd),
                                             _p(d_n), _p(d_l), _stream()))
        return d_kd, d_n, d_l


def shade_directional(kd, normal, light_params):
    """kd, normal [B,H,W,3] (kd may be a channel slice of the texture field's output), light_params [B|1,5] =
    (dir.xyz, ambient, diffuse) -> (shaded [B,H,W,3], shading [B,H,W,1]); one kernel per direction."""
    return _ShadeDirectional.apply(kd, normal, light_params)


class _AnalyticField(torch.autograd.Function):
    """Benchmark stand-in for the field MLPs (M1a; not a reference interface): act(x W), one kernel per
    direction (csrc/analytic_field.cu)."""

    @staticmethod
    def forward(ctx, x, weight, squash):
        x = _f32(x, "x"); weight = _f32(weight, "weight")
        C = weight.shape[1]
        N = x.numel() // 3
        out = torch.empty(*x.shape[:-1], 3 * C if squash else C, device=x.device)
        _call("b2a_analytic_field_fwd", (_p(x), _p(weight), C, int(squash), N, _p(out), _stream()))
        ctx.save_for_backward(x, weight)
        ctx.squash = bool(squash)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = _f32(g, "d_out")
        d_x = torch.empty_like(x)
        _call("b2a_analytic_field_bwd", (_p(x), _p(weight), weight.shape[1], int(ctx.squash), x.numel() // 3, _p(g), _p(d_x), _stream()))
        return d_x, None, None


def analytic_field(x, weight, squash):
    """x [...,3], weight [3,C] -> sin(x W) [...,C] (squash False) or cat([sigmoid(x W)] * 3) [...,3C] (squash True)."""
    return _AnalyticField.apply(x, weight, bool(squash))


# ---------------------------------------------------------------------------------------------------------------
# Fused geometry half of render_mesh: clip transform -> rasterize -> g-buffer (-> antialias analysis) as ONE autograd
# node.  Same kernels as the separate ops above; what it removes is host work per step (three autograd nodes forward and
# backward, two gradient-sum launches, two zero fills): the hot path at B=16 is host-bound.
# ---------------------------------------------------------------------------------------------------------------
class _RenderGeometry(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v_pos, v_nrm, prior_pos, mtx, w2c, campos, tri, opp, H, W, spp, two_sided, want, need_aa):
        L = _L()
        v_pos = _f32(v_pos, "v_pos"); v_nrm = _f32(v_nrm, "v_nrm"); prior_pos = _f32(prior_pos, "prior_pos")
        mtx = _f32(mtx, "matrix"); w2c = _f32(w2c, "w2c"); campos = _f32(campos, "campos")
        B, V = mtx.shape[0], v_pos.shape[1]
        Bp, Bq, F = v_pos.shape[0], prior_pos.shape[0], tri.shape[0]
        if Bp != B or w2c.shape != (B, 4, 4) or campos.shape != (B, 3) or v_nrm.shape != v_pos.shape:
            raise _lib.B2AError("render geometry: inconsistent batch shapes")
        dev = v_pos.device
        st = _stream()
        fH, fW = H * spp, W * spp
        clip = torch.empty(B, V, 4, device=dev)
        ws = _workspace(_size(L.b2a_rasterize_workspace_bytes, B, F, fH, fW), dev)
        rast = torch.empty(B, fH, fW, 4, device=dev)
        use_cov = spp == 1
        cov_list = torch.empty((B * fH * fW, 4), dtype=_i32, device=dev) if use_cov else None
        cov_count = torch.empty(1, dtype=_i32, device=dev) if use_cov else None
        outs = [torch.empty(B, H, W, 3, device=dev) if k in want else None for k in GB_KEYS]
        packed = _workspace(_size(L.b2a_gbuffer_pack_bytes, B, Bq, V), dev)
        aa_ctx = None
        if need_aa and (fH * fW) % 32 == 0 and F < (1 << 28):
            aa_ctx = _workspace(_size(L.b2a_antialias_workspace_bytes, B, fH, fW), dev)
        # clip transform -> rasterize -> g-buffer -> antialias analysis: ONE C-ABI call (b2a_render_geometry_fwd)
        _call("b2a_render_geometry_fwd", (_p(v_pos), _p(v_nrm), _p(prior_pos), Bq, _p(mtx), _p(w2c), _p(campos), _p(tri), _p(opp), int(two_sided), B, V, F,
                                          H, W, spp, _p(ws), ws.numel(), _p(packed), packed.numel(), _p(clip), _p(rast), _p(cov_list), _p(cov_count),
                                          *[_p(o) for o in outs], _p(aa_ctx), 0 if aa_ctx is None else aa_ctx.numel(), st),
              launches=8 if aa_ctx is not None else 6)
        ctx.save_for_backward(rast, clip, tri, v_pos, v_nrm, prior_pos, mtx, w2c, campos, cov_list, cov_count, packed)
        ctx.cfg = (spp, int(two_sided), H, W, tuple(o is not None for o in outs))
        ctx.set_materialize_grads(False)   # an unused output (rast outside 'flow' mode) must arrive as None, not as a zero tensor
        aa_out = aa_ctx if aa_ctx is not None else torch.empty(0, dtype=torch.uint8, device=dev)
        ctx.mark_non_differentiable(aa_out)
        return (clip, rast, aa_out) + tuple(o for o in outs if o is not None)     # only the requested g-buffers, in GB_KEYS order

    @staticmethod
    def backward(ctx, d_clip_up, d_rast, _aa, *grads):
        L = _L()
        rast, clip, tri, v_pos, v_nrm, prior_pos, mtx, w2c, campos, cov_list, cov_count, packed = ctx.saved_tensors
        spp, two_sided, H, W, present = ctx.cfg
        B, V, F = v_pos.shape[0], v_pos.shape[1], tri.shape[0]
        st = _stream()
        need = ctx.needs_input_grad
        it = iter(grads)                                                      # grads arrive for the present outputs only
        gs = [next(it) if p else None for p in present]
        gs = [(_f32(g, "d_gb") if g is not None else None) for g in gs]
        have_gb = any(g is not None for g in gs)
        d_v_pos = torch.empty_like(v_pos)
        d_v_nrm = torch.empty_like(v_nrm) if need[1] else None
        d_prior = torch.empty_like(prior_pos) if need[2] else None
        d_mtx = torch.zeros_like(mtx) if need[3] else None
        d_w2c = torch.zeros_like(w2c) if need[4] else None
        d_campos = torch.zeros_like(campos) if need[5] else None
        up = _f32(d_clip_up, "d_clip") if d_clip_up is not None else None
        if d_rast is not None:      # only the 'flow' mode interpolates with a differentiable rast (render.py:281-288)
            d_clip = torch.zeros_like(clip)
            _call("b2a_rasterize_bwd", (_p(clip), _p(tri), _p(rast), _p(_f32(d_rast, "d_rast")), B, V, F, rast.shape[1], rast.shape[2],
                                              _p(d_clip), st))
            up = d_clip if up is None else up + d_clip
        # g-buffer / rasterize adjoint over the covered-pixel list, then ONE per-vertex pass: accumulator rows + antialias clip
        # gradient -> clip-transform adjoint -> d_v_pos / d_v_nrm / d_prior (b2a_render_geometry_bwd)
        acc = _gb_accumulator(_size(L.b2a_gbuffer_bwd_workspace_bytes, B, V), rast.device)
        _call("b2a_render_geometry_bwd", (_p(rast), spp, _p(mtx), _p(tri), _p(v_pos), _p(v_nrm), _p(prior_pos), prior_pos.shape[0], _p(w2c), _p(campos),
                                          two_sided, B, V, F, H, W, _p(packed), packed.numel(), _p(cov_list), _p(cov_count), *[_p(g) for g in gs], _p(up),
                                          _p(acc), acc.numel(), 1, _p(d_v_pos), _p(d_v_nrm), _p(d_prior), _p(d_mtx), _p(d_w2c), _p(d_campos), st),
              launches=2 if have_gb else 1)
        return (d_v_pos if need[0] else None, d_v_nrm, d_prior, d_mtx, d_w2c, d_campos, None, None, None, None, None, None, None, None)


def render_geometry(v_pos, v_nrm, prior_pos, mtx, w2c, campos, tri, opp, resolution, spp=1, two_sided=True, want=("cam_nrm", "tex_pos"),
                    need_aa=True):
    """-> (v_pos_clip [B,V,4], rast [B,H*spp,W*spp,4], aa_ctx | None, dict of g-buffers [B,H,W,3]).  rast is differentiable
    (barycentric gradients of a later ops.interpolate flow into the clip positions, as with nvdiffrast)."""
    cpp = _cpp()
    if cpp is not None:
        mask = sum(1 << i for i, k in enumerate(GB_KEYS) if k in want)
        outs = cpp.render_geometry(v_pos, v_nrm, prior_pos, mtx, w2c, campos, _idx32(tri, "tri"), _idx32(opp, "opp"), int(resolution[0]),
                                   int(resolution[1]), int(spp), bool(two_sided), mask, bool(need_aa))
    else:
        outs = _RenderGeometry.apply(v_pos, v_nrm, prior_pos, mtx, w2c, campos, _idx32(tri, "tri"), _idx32(opp, "opp"), int(resolution[0]),
                                     int(resolution[1]), int(spp), bool(two_sided), tuple(want), bool(need_aa))
    clip, rast, aa = outs[0], outs[1], outs[2]
    return clip, rast, (aa if aa.numel() else None), dict(zip([k for k in GB_KEYS if k in want], outs[3:]))
