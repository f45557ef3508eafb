"""Drop-in for the reference's model/geometry/dmtet.py: DMTet, DMTetGeometry, sdf_bce_reg_loss.

Same constructor arguments, attributes and return types; the extraction itself runs in the
sm_100a kernels of libb2a.so (csrc/marching_tets.cu) through `ops.marching_tets` - no torch.unique, no boolean-mask
gathers, ONE device->host read (the output sizes).  The SDF MLP stays a PyTorch module (out of scope, §8a R1).
"""
import math
import os

import numpy as np
import torch

from .. import field_mlp
from .. import ops
from ..render import mesh


class DMTet:
    """Differentiable marching tetrahedra with the reference's call signature (dmtet.py:104-155):
    `(pos_nx3, sdf_n, tet_fx4) -> (verts [V,3] f32, faces [F,3] i64, uvs [4N^2,2] f32, uv_idx [F,3] i64)`."""

    def __init__(self, device="cuda"):
        self.device = device
        self._grid = None
        self._grid_key = None
        self._grid_src = None
        self._uv_cache = {}

    def to(self, device):
        self.device = device
        return self

    def grid_for(self, tet_fx4, num_verts):
        """Static edge tables for a tet index tensor; cached, so `load_tets` pays the table build once per grid."""
        key = (tuple(tet_fx4.shape), int(num_verts), tet_fx4._version)
        # the keyed tensor itself is held (compared with `is`): its address cannot be recycled for another grid while cached
        if self._grid_src is not tet_fx4 or self._grid_key != key:
            self._grid = ops.TetGrid(tet_fx4, num_verts)
            self._grid_key = key
            self._grid_src = tet_fx4
        return self._grid

    def uv_table(self, num_tets, device):
        """The per-tet UV atlas of map_uv (dmtet.py:69-84): a constant of the grid size, so it is built once and
        shared (the reference rewrites these 4N^2 rows - 50 MB at res 64 - on every call)."""
        N = int(math.ceil(math.sqrt((num_tets * 2 + 1) // 2)))
        key = (N, str(device))
        if key not in self._uv_cache:
            lin = torch.linspace(0, 1 - (1 / N), N, dtype=torch.float32, device=device)
            tex_y, tex_x = torch.meshgrid(lin, lin, indexing="ij")
            pad = 0.9 / N
            uvs = torch.stack([tex_x, tex_y, tex_x + pad, tex_y, tex_x + pad, tex_y + pad, tex_x, tex_y + pad], dim=-1)
            self._uv_cache = {key: uvs.view(-1, 2)}
        return self._uv_cache[key]

    def extract(self, pos_nx3, sdf_n, grid):
        """Fast path used by DMTetGeometry: returns (verts, faces_i64, uv_idx_i64, faces_i32)."""
        verts, faces, uv_idx, faces32, _ = ops.marching_tets(pos_nx3, sdf_n, grid)
        return verts, faces, uv_idx, faces32

    def __call__(self, pos_nx3, sdf_n, tet_fx4):
        grid = self.grid_for(tet_fx4, pos_nx3.shape[0])
        verts, faces, uv_idx, _ = self.extract(pos_nx3, sdf_n, grid)
        return verts, faces, self.uv_table(grid.T, pos_nx3.device), uv_idx


def sdf_bce_reg_loss(sdf, all_edges):
    """Reference dmtet.py:161-169 (a regulariser over the static edge list; plain PyTorch, off the render path)."""
    pairs = sdf[all_edges.reshape(-1)].reshape(-1, 2)
    crossing = torch.sign(pairs[..., 0]) != torch.sign(pairs[..., 1])
    pairs = pairs[crossing]
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    return bce(pairs[..., 0], (pairs[..., 1] > 0).float()) + bce(pairs[..., 1], (pairs[..., 0] > 0).float())


def _field_networks():
    """CoordMLP / CoordMLP_Mod: the reference's own classes when this package is overlaid on the reference tree
    (model.networks importable), else the package's minimal equivalents used by the standalone bench and tests."""
    try:
        from model.networks import CoordMLP, CoordMLP_Mod  # noqa: the reference's PyTorch modules, unchanged
        return CoordMLP, CoordMLP_Mod
    except Exception:
        from ..networks import CoordMLP, CoordMLP_Mod
        return CoordMLP, CoordMLP_Mod


class DMTetGeometry(torch.nn.Module):
    """Reference dmtet.py:175-310.  Attributes kept: verts, indices, grid_res, grid_scale, all_edges, current_sdf,
    mesh_verts, mlp (parameter names unchanged, so checkpoints load)."""

    def __init__(self, grid_res, spatial_scale, num_layers=None, hidden_size=None, embedder_freq=None, embed_concat_pts=True,
                 init_sdf=None, jitter_grid=0., symmetrize=False, condition_choice=None, **kwargs):
        super().__init__()
        self.grid_res = grid_res
        self.marching_tets = DMTet()
        self.grid_scale = spatial_scale
        self.init_sdf = init_sdf
        self.jitter_grid = jitter_grid
        self.symmetrize = symmetrize
        self.tets_root = kwargs.get("tets_root", "data/tets")
        self.synthetic_tets = bool(kwargs.get("synthetic_tets", False))
        # Narrow-band SDF evaluation (off by default = the reference's full-grid evaluation): `narrow_band = (k, M)`
        # evaluates the SDF network only on grid vertices within k grid edges of the surface found by the last FULL evaluation, and
        # refreshes the full grid every M calls (and whenever the surface reaches the rim of the band).  Extraction, the BCE
        # regulariser and every gradient only involve vertices on sign-changing edges, so inside the band the results are those
        # of the full evaluation.
        self.narrow_band = kwargs.get("narrow_band", None)
        self._band = None
        self.load_tets(self.grid_res, self.grid_scale)
        CoordMLP, CoordMLP_Mod = _field_networks()
        embedder_scalar = 2 * np.pi / self.grid_scale * 0.9
        common = dict(dropout=0, activation=None, min_max=None, n_harmonic_functions=embedder_freq,
                      embedder_scalar=embedder_scalar, embed_concat_pts=embed_concat_pts)
        if condition_choice == "mod":
            self.mlp = CoordMLP_Mod(3, 1, num_layers, nf=hidden_size, condition_dim=128, **common)
        else:
            self.mlp = CoordMLP(3, 1, num_layers, nf=hidden_size, **common)

    def load_tets(self, grid_res=None, scale=None):
        if grid_res is None:
            grid_res = self.grid_res
        else:
            self.grid_res = grid_res
        if scale is None:
            scale = self.grid_scale
        else:
            self.grid_scale = scale
        path = os.path.join(self.tets_root, "{}_tets.npz".format(grid_res))
        if not os.path.isfile(path):
            # The reference raises here (np.load, dmtet.py:223): a different grid would silently extract a different mesh from
            # a reference checkpoint.  Offline benches / tests opt in to a synthetic Kuhn grid in the reference's schema
            # (`synthetic_tets=True` or B2A_SYNTHETIC_TETS=1); it is written atomically so concurrent ranks never read a torn file.
            if not (self.synthetic_tets or os.environ.get("B2A_SYNTHETIC_TETS", "0") == "1"):
                raise FileNotFoundError("tet grid %s not found (run data/tets/download_tets.sh; or pass synthetic_tets=True / "
                                        "B2A_SYNTHETIC_TETS=1 for a synthetic Kuhn grid)" % path)
            from ..synthetic import write_tet_npz
            path = write_tet_npz(grid_res, self.tets_root)
        tets = np.load(path)
        self.verts = torch.tensor(tets["vertices"], dtype=torch.float32, device="cuda") * scale
        self.indices = torch.tensor(tets["indices"], dtype=torch.long, device="cuda")
        self.generate_edges()

    def generate_edges(self):
        with torch.no_grad():
            self.grid = self.marching_tets.grid_for(self.indices, self.verts.shape[0])
            self._all_edges = None

    @property
    def all_edges(self):
        if self._all_edges is None:
            self._all_edges = self.grid.all_edges()
        return self._all_edges

    def get_sdf(self, pts=None, total_iter=0, feats=None):
        if pts is None:
            pts = self.verts
        if self.symmetrize:
            xs, ys, zs = pts.unbind(-1)
            pts = torch.stack([xs.abs(), ys, zs], -1)
        if feats is not None:
            feats = feats.unsqueeze(0).repeat(pts.shape[0], 1)
        if feats is None and pts.dim() == 2 and pts.is_cuda and field_mlp.supported(self.mlp, pts, None):
            # the SDF network is a CoordMLP too: the same tcgen05 GEMMs as the texture / DINO fields (csrc/field_mlp.cu); a
            # double backward (the eikonal regulariser differentiates get_sdf's gradient) stays on PyTorch's ops
            sdf = field_mlp.coord_mlp_rows(self.mlp, pts, None, None, 1) if not torch.is_grad_enabled() or not pts.requires_grad else self.mlp(pts, feat=feats)
        else:
            sdf = self.mlp(pts, feat=feats)
        if self.init_sdf is None:
            pass
        elif type(self.init_sdf) in [float, int]:
            sdf = sdf + self.init_sdf
        elif self.init_sdf == "sphere":
            sdf = sdf + (self.grid_scale * 0.25 - pts.norm(dim=-1, keepdim=True))
        elif self.init_sdf == "ellipsoid":
            xs, ys, zs = pts.unbind(-1)
            sdf = sdf + (self.grid_scale * 0.15 - torch.stack([xs, ys, zs / 2], -1).norm(dim=-1, keepdim=True))
        else:
            raise NotImplementedError
        return sdf

    def get_sdf_gradient(self, feats=None):
        num_samples = 5000
        sample_points = (torch.rand(num_samples, 3, device=self.verts.device) - 0.5) * self.grid_scale
        mesh_verts = self.mesh_verts.detach() + (torch.rand_like(self.mesh_verts) - 0.5) * 0.1 * self.grid_scale
        rand_idx = torch.randperm(len(mesh_verts), device=mesh_verts.device)[:5000]
        sample_points = torch.cat([sample_points, mesh_verts[rand_idx]], 0)
        sample_points.requires_grad = True
        y = self.get_sdf(pts=sample_points, feats=feats)
        try:
            return torch.autograd.grad(outputs=[y], inputs=sample_points, grad_outputs=torch.ones_like(y), create_graph=True,
                                       retain_graph=True, only_inputs=True)[0]
        except RuntimeError:  # validation runs under no_grad
            return torch.zeros_like(sample_points)

    def get_sdf_reg_loss(self, feats=None):
        return {"sdf_bce_reg_loss": sdf_bce_reg_loss(self.current_sdf, self.all_edges).mean(),
                "sdf_gradient_reg_loss": ((self.get_sdf_gradient(feats=feats).norm(dim=-1) - 1) ** 2).mean()}

    @torch.no_grad()
    def getAABB(self):
        return torch.min(self.verts, dim=0).values, torch.max(self.verts, dim=0).values

    def _sdf_full_and_band(self, v_deformed, total_iter, feats, k):
        """Full-grid evaluation + the band for the next calls: vertices within k grid edges of a sign-changing edge (dilation over
        the static edge list), and its outermost ring (the guard: a sign change there means the surface has reached the rim)."""
        sdf = self.get_sdf(v_deformed, total_iter=total_iter, feats=feats)
        with torch.no_grad():
            e = self.all_edges
            a, b = e[:, 0], e[:, 1]
            occ = sdf.detach().reshape(-1) > 0
            band = torch.zeros_like(occ)
            cross = occ[a] != occ[b]
            band[a[cross]] = True
            band[b[cross]] = True
            ring = band
            for _ in range(int(k)):
                grow = band[a] | band[b]
                nxt = band.clone()
                nxt[a[grow]] = True
                nxt[b[grow]] = True
                ring = nxt & ~band
                band = nxt
            idx = torch.nonzero(band).squeeze(1)
            self._band = dict(idx=idx, ring=ring[idx], sdf=sdf.detach().clone(), occ_ring=occ[idx][ring[idx]], age=0,
                              flag=torch.zeros(1, dtype=torch.int32).pin_memory(), rows=int(idx.numel()))
        return sdf

    def _sdf_narrow_band(self, v_deformed, total_iter, feats):
        """SDF values for the extraction with the network evaluated on the band only; everything else keeps the values (hence the
        signs) of the last full evaluation.  Returns None when the band must be refreshed."""
        k, every = self.narrow_band
        bd = self._band
        if bd is None or bd["age"] + 1 >= int(every) or bd["sdf"].shape[0] != v_deformed.shape[0]:
            return None
        idx = bd["idx"]
        sdf_band = self.get_sdf(v_deformed.index_select(0, idx), total_iter=total_iter, feats=feats)
        # guard: the sign pattern on the band's outermost ring must be unchanged (device-side count into pinned memory, read
        # together with the extraction's own size hand-off - no extra synchronisation)
        moved = ((sdf_band.detach().reshape(-1)[bd["ring"]] > 0) != bd["occ_ring"]).sum().to(torch.int32)
        bd["flag"].copy_(moved.reshape(1), non_blocking=True)
        bd["age"] += 1
        return bd["sdf"].index_copy(0, idx, sdf_band)

    def getMesh(self, material=None, total_iter=0, jitter_grid=True, feats=None):
        v_deformed = self.verts
        if jitter_grid and self.jitter_grid > 0:
            jitter = (torch.rand(1, device=v_deformed.device) * 2 - 1) * self.jitter_grid * self.grid_scale
            v_deformed = v_deformed + jitter
        self.sdf_rows_evaluated = v_deformed.shape[0]
        if self.narrow_band:
            sdf = self._sdf_narrow_band(v_deformed, total_iter, feats)
            if sdf is not None:
                self.current_sdf = sdf
                verts, faces, uv_idx, faces32 = self.marching_tets.extract(v_deformed, self.current_sdf, self.grid)
                if int(self._band["flag"].item()) == 0:         # (the extraction has synchronised: the flag is final)
                    self.sdf_rows_evaluated = self._band["rows"]
                    self.mesh_verts = verts
                    uvs = self.marching_tets.uv_table(self.grid.T, verts.device)
                    return mesh.make_mesh(verts[None], faces[None], uvs[None], uv_idx[None], material, faces_i32=faces32)
            self.current_sdf = self._sdf_full_and_band(v_deformed, total_iter, feats, self.narrow_band[0])
        else:
            self.current_sdf = self.get_sdf(v_deformed, total_iter=total_iter, feats=feats)
        verts, faces, uv_idx, faces32 = self.marching_tets.extract(v_deformed, self.current_sdf, self.grid)
        self.mesh_verts = verts
        uvs = self.marching_tets.uv_table(self.grid.T, verts.device)
        return mesh.make_mesh(verts[None], faces[None], uvs[None], uv_idx[None], material, faces_i32=faces32)
