"""`nvdiffrast.torch`-compatible surface over libb2a.so, so the reference files that import nvdiffrast directly
(model/models/AnimalModel.py:9,236; model/render/material.py:13,116; visualization/visualize_results.py:225) run
unchanged.  Only what those call sites use is provided: contexts, rasterize, DepthPeeler (first layer),
interpolate, antialias.  `texture` belongs to out-of-scope callers (EnvironmentLight, Texture2D) and raises.
"""
from .. import ops


class RasterizeCudaContext:
    def __init__(self, device=None):
        self.device = device


class RasterizeGLContext(RasterizeCudaContext):
    def __init__(self, output_db=True, mode="automatic", device=None):
        super().__init__(device)


def rasterize(glctx, pos, tri, resolution, ranges=None, grad_db=True):
    """-> (rast [B,H,W,4] = (u, v, z/w, triangle_id+1), rast_db).  The image-space derivative buffer is dead in the
    reference (every dr.interpolate call passes rast_db=None, render.py:182-209), so a zero tensor is returned."""
    if ranges is not None:
        raise NotImplementedError("range mode is not used by the reference")
    rast = ops.rasterize(pos, tri, resolution)
    return rast, rast.new_zeros(rast.shape)


class DepthPeeler:
    def __init__(self, glctx, pos, tri, resolution):
        self.pos, self.tri, self.resolution = pos, tri, resolution
        self.layer = 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def rasterize_next_layer(self):
        if self.layer > 0:
            raise NotImplementedError("only the first depth layer is ever requested (num_layers=1, AnimalModel.py:247)")
        self.layer += 1
        return rasterize(None, self.pos, self.tri, self.resolution)


def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    if rast_db is not None and diff_attrs is not None:
        raise NotImplementedError("attribute pixel differentials are not used by the reference")
    return ops.interpolate(attr, rast, tri), None


def antialias(color, rast, pos, tri, topology_hash=None, pos_gradient_boost=1.0):
    if pos_gradient_boost != 1.0:
        raise NotImplementedError("pos_gradient_boost != 1 is not used by the reference")
    return ops.antialias(color, rast, pos, tri)


def texture(*args, **kwargs):
    raise NotImplementedError("nvdiffrast.torch.texture is only reached from out-of-scope callers (SURVEY.md §8c)")
