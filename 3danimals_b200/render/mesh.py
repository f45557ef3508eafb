"""Drop-in for the reference's model/render/mesh.py: Mesh, make_mesh, auto_normals (+ the edge utilities).

`auto_normals` runs in libb2a.so (csrc/normals.cu, one launch for the whole batch, analytic backward) - when somebody
reads `mesh.v_nrm`: make_mesh / auto_normals only mark the normals as pending.  On the training path the prior shape's
normals are never read (only the posed instances' normals reach the renderer, render.py:195), so a third of the
reference's normal passes (mesh.py:276-304 runs three times per step) is not executed at all.  Tangents are
numerically dead on every path of the reference (the perturbed normal is the constant (0,0,1), so the
shading normal never depends on them) and the 4N^2-row UV atlas they are derived from is a per-grid constant; both
are therefore materialised lazily - `mesh.v_tng` / `mesh.v_tex` have the reference's values when somebody reads them,
and cost nothing on the training path.
"""
import torch

from .. import ops


def _dot(x, y):
    return torch.sum(x * y, -1, keepdim=True)


def _safe_normalize(x, eps=1e-20):
    return x / torch.sqrt(torch.clamp(_dot(x, x), min=eps))


class Mesh:
    """Minibatched mesh with shared connectivity (reference mesh.py:21-175).  Same attribute names."""

    def __init__(self, v_pos=None, t_pos_idx=None, v_nrm=None, t_nrm_idx=None, v_tex=None, t_tex_idx=None, v_tng=None,
                 t_tng_idx=None, material=None, base=None):
        self.v_pos = v_pos
        self._v_nrm = v_nrm
        self._nrm_pending = False   # auto_normals: compute v_nrm from v_pos on first access
        self._v_tex = v_tex
        self._v_tng = v_tng
        self.t_pos_idx = t_pos_idx
        self.t_nrm_idx = t_nrm_idx
        self.t_tex_idx = t_tex_idx
        self._t_tng_idx = t_tng_idx
        self.material = material
        self.t_pos_idx_i32 = None   # [F,3] int32 copy of t_pos_idx[0] for the kernels
        self._opp = None            # edge adjacency of the topology (antialiasing), built on first render
        if base is not None:
            self.copy_none(base)

    # -- lazily materialised attributes --------------------------------------------------------------------------
    @property
    def v_nrm(self):
        if self._nrm_pending:
            self._nrm_pending = False
            self._v_nrm = ops.vertex_normals(self.v_pos, self.tri_i32())
            if torch.is_anomaly_enabled():
                assert torch.all(torch.isfinite(self._v_nrm))
        return self._v_nrm

    @v_nrm.setter
    def v_nrm(self, v):
        self._v_nrm = v
        self._nrm_pending = False

    @property
    def v_tex(self):
        v = self._v_tex
        if v is not None and self.v_pos is not None and v.shape[0] != self.v_pos.shape[0] and v.shape[0] == 1:
            v = v.expand(self.v_pos.shape[0], -1, -1)  # what `.repeat(B,1,1)` holds, without the copy
        return v

    @v_tex.setter
    def v_tex(self, v):
        self._v_tex = v

    @property
    def v_tng(self):
        if self._v_tng is None and (self._v_nrm is not None or self._nrm_pending) and self._v_tex is not None and self.t_tex_idx is not None:
            self._v_tng = _tangents(self)
        return self._v_tng

    @v_tng.setter
    def v_tng(self, v):
        self._v_tng = v

    @property
    def t_tng_idx(self):
        return self._t_tng_idx if self._t_tng_idx is not None else self.t_nrm_idx

    @t_tng_idx.setter
    def t_tng_idx(self, v):
        self._t_tng_idx = v

    def tri_i32(self):
        if self.t_pos_idx_i32 is None:
            self.t_pos_idx_i32 = self.t_pos_idx[0].to(torch.int32).contiguous()
        return self.t_pos_idx_i32

    def edge_adjacency(self):
        if self._opp is None:
            self._opp = ops.edge_adjacency(self.tri_i32(), self.v_pos.shape[1])
        return self._opp

    # -- reference API -------------------------------------------------------------------------------------------
    def __len__(self):
        return len(self.v_pos)

    def copy_none(self, other):
        for name in ("v_pos", "t_pos_idx", "t_nrm_idx", "_v_tex", "t_tex_idx", "_v_tng", "_t_tng_idx", "material"):
            if getattr(self, name) is None:
                setattr(self, name, getattr(other, name))
        if self._v_nrm is None and not self._nrm_pending:      # normals: take the other's, computed or still pending
            if other._nrm_pending and self.v_pos is other.v_pos and self.t_pos_idx is other.t_pos_idx:
                self._nrm_pending = True
            else:
                self._v_nrm = other.v_nrm
        if self.t_pos_idx is other.t_pos_idx:
            if self.t_pos_idx_i32 is None:
                self.t_pos_idx_i32 = other.t_pos_idx_i32
            if self._opp is None:
                self._opp = other._opp

    def clone(self):
        out = Mesh(base=self)
        if out._nrm_pending:          # a clone holds values, not a recipe
            out.v_nrm = self.v_nrm
        for name in ("v_pos", "t_pos_idx", "_v_nrm", "t_nrm_idx", "_v_tex", "t_tex_idx", "_v_tng", "_t_tng_idx"):
            v = getattr(out, name)
            if v is not None:
                setattr(out, name, v.clone().detach())
        return out

    def detach(self):
        return self.clone()

    def _same_topology(self, verts, uvs):
        m = make_mesh(verts, self.t_pos_idx, uvs, self.t_tex_idx, self.material, faces_i32=self.tri_i32())
        m._opp = self._opp
        return m

    def _uv_rows(self):
        return self._v_tex[:1] if self._v_tex is not None else None

    def extend(self, N):
        return self._same_topology(self.v_pos.repeat(N, 1, 1), self._expand_uv(self.v_pos.shape[0] * N))

    def deform(self, deformation):
        assert deformation.shape[1] == self.v_pos.shape[1] and deformation.shape[2] == 3
        verts = self.v_pos + deformation
        return self._same_topology(verts, self._expand_uv(len(verts)))

    def _expand_uv(self, n):
        uv = self._v_tex
        if uv is None:
            return None
        return uv[:1].expand(n, -1, -1) if uv.shape[0] != n else uv

    def get_m_to_n(self, m, n):
        verts = self.v_pos[m:n, ...]
        return self._same_topology(verts, self._expand_uv(verts.shape[0]))

    def first_n(self, n):
        return self.get_m_to_n(0, n)

    def get_n(self, n):
        return self.get_m_to_n(n, n + 1)


def load_mesh(filename, mtl_override=None):
    """mesh.py:181-185.  `load_obj` is the reference's own reader, reachable in overlay mode (render/obj.py re-exports it)."""
    import os
    from . import obj
    if os.path.splitext(filename)[1] == ".obj":
        return obj.load_obj(filename, clear_ks=True, mtl_override=mtl_override)
    assert False, "Invalid mesh file extension"


def aabb(mesh):
    return torch.min(mesh.v_pos, dim=0).values, torch.max(mesh.v_pos, dim=0).values


def _sorted_edges(attr_idx):
    idx = attr_idx[0]
    e = torch.stack((idx[:, [0, 1]], idx[:, [1, 2]], idx[:, [2, 0]]), dim=1).reshape(-1, 2)
    order = e[:, 0] > e[:, 1]
    return torch.stack((e.min(1).values, e.max(1).values), -1), order


def compute_edges(attr_idx, return_inverse=False):
    """Unique (min,max) edges of a triangle index list (reference mesh.py:196-214; used by the regularisers)."""
    with torch.no_grad():
        return torch.unique(_sorted_edges(attr_idx)[0], dim=0, return_inverse=return_inverse)


def compute_edge_to_face_mapping(attr_idx, return_inverse=False):
    """Reference mesh.py:219-250: per unique edge, the triangle seeing it in ascending / descending vertex order."""
    with torch.no_grad():
        edges, order = _sorted_edges(attr_idx)
        unique_edges, idx_map = torch.unique(edges, dim=0, return_inverse=True)
        tris = torch.arange(attr_idx[0].shape[0], device=edges.device).repeat_interleave(3)
        tris_per_edge = torch.zeros((unique_edges.shape[0], 2), dtype=torch.int64, device=edges.device)
        tris_per_edge[idx_map[~order], 0] = tris[~order]
        tris_per_edge[idx_map[order], 1] = tris[order]
        return tris_per_edge


def unit_size(mesh):
    with torch.no_grad():
        vmin, vmax = aabb(mesh)
        scale = 2 / torch.max(vmax - vmin).item()
        return Mesh((mesh.v_pos - (vmax + vmin) / 2) * scale, base=mesh)


def center_by_reference(base_mesh, ref_aabb, scale):
    center = (ref_aabb[0] + ref_aabb[1]) * 0.5
    scale = scale / torch.max(ref_aabb[1] - ref_aabb[0]).item()
    return Mesh((base_mesh.v_pos - center[None, ...]) * scale, base=base_mesh)


def auto_normals(imesh):
    """Smooth area-weighted vertex normals (reference mesh.py:276-304) in one fused launch + analytic backward."""
    if not imesh.v_pos.is_cuda:
        raise ops._lib.B2AError("auto_normals needs CUDA tensors (the B200 hot path has no CPU fallback)")
    out = Mesh(t_nrm_idx=imesh.t_pos_idx)
    out._nrm_pending = True        # evaluated by the v_nrm property on first read (one fused launch + analytic backward)
    out.copy_none(imesh)
    return out


def _tangents(imesh):
    """Reference mesh.py:310-350 (MikkTSpace-style), evaluated only when a caller actually reads `v_tng`."""
    B = imesh.v_pos.shape[0]
    pos = [imesh.v_pos[:, imesh.t_pos_idx[0, :, i]] for i in range(3)]
    uv = imesh._v_tex            # [B,Nuv,2] per-image texcoords, or the shared [1,Nuv,2] atlas (broadcasts below)
    tex = [uv[:, imesh.t_tex_idx[0, :, i]] for i in range(3)]
    uve1, uve2 = tex[1] - tex[0], tex[2] - tex[0]
    pe1, pe2 = pos[1] - pos[0], pos[2] - pos[0]
    nom = pe1 * uve2[..., 1:2] - pe2 * uve1[..., 1:2]
    denom = uve1[..., 0:1] * uve2[..., 1:2] - uve1[..., 1:2] * uve2[..., 0:1]
    tang = nom / torch.where(denom > 0.0, torch.clamp(denom, min=1e-6), torch.clamp(denom, max=-1e-6))
    tangents = torch.zeros_like(imesh.v_nrm)
    tansum = torch.zeros_like(imesh.v_nrm)
    for i in range(3):
        idx = imesh.t_nrm_idx[..., i:i + 1].expand(B, -1, 3)
        tangents = tangents.scatter_add(1, idx, tang)
        tansum = tansum.scatter_add(1, idx, torch.ones_like(tang))
    tangents = _safe_normalize(tangents / tansum)
    return _safe_normalize(tangents - _dot(tangents, imesh.v_nrm) * imesh.v_nrm)


def compute_tangents(imesh):
    return Mesh(v_tng=_tangents(imesh), t_tng_idx=imesh.t_nrm_idx, base=imesh)


def make_mesh(verts, faces, uvs, uv_idx, material, faces_i32=None):
    """verts [B,V,3], faces [1,F,3], uvs [B|1,Nuv,2], uv_idx [1,F,3] -> Mesh with normals (reference mesh.py:355-375)."""
    assert len(verts.shape) == 3 and len(faces.shape) == 3 and (uvs is None or len(uvs.shape) == 3) and \
        (uv_idx is None or len(uv_idx.shape) == 3), "All components must be batched."
    assert faces.shape[0] == 1 and (uv_idx is None or uv_idx.shape[0] == 1), "Every mesh must share the same edge connectivity."
    assert uvs is None or verts.shape[0] == uvs.shape[0] or uvs.shape[0] == 1, "Batch size must be consistent."
    ret = Mesh(verts, faces, v_tex=uvs, t_tex_idx=uv_idx, material=material)
    ret.t_pos_idx_i32 = faces_i32
    return auto_normals(ret)
