"""Drop-in for the reference's model/render/obj.py: `write_obj` (obj.py:128-177) - the step AFTER the hot path in the test /
visualize configs, called once per exported instance by `misc.save_obj` (model/utils/misc.py:187).

The reference writes the file with four Python loops, one `f.write` per vertex / texcoord / normal / face (2.3 s for a
configs[4]-sized mesh).  Here the whole text is produced by ONE C-ABI call (`b2a_obj_format`, csrc/obj_format.cu: all host
threads, shortest-round-trip float formatting) and written with one `write`; the bytes are identical, including the
reference's number format ('{}'.format(np.float32) = repr of the value widened to double; texcoord v flipped in float32).
The material half (`material.save_mtl`, material.py:106-140) stays the reference's own function: it renders the texture maps
through `render_uv` / `nvdiffrast.torch`, which under the overlay are this package's.

`load_obj` and every other name of the reference module are re-exported from the reference's own file when the reference
tree is importable (overlay mode); standalone, only `write_obj` exists.  No CPU fallback: without libb2a.so the call raises.
"""
import ctypes as C
import importlib
import importlib.util
import os
import sys

import numpy as np

from .. import _lib

_ref_cache = {}


def _reference_obj():
    """The reference's own model/render/obj.py under a private name (overlay mode), or None."""
    pkg = sys.modules.get("model.render")
    for d in getattr(pkg, "__path__", None) or []:
        f = os.path.join(d, "obj.py")
        if f in _ref_cache:
            return _ref_cache[f]
        if os.path.isfile(f):
            try:
                spec = importlib.util.spec_from_file_location("model.render._reference_obj", f)
                mod = importlib.util.module_from_spec(spec)
                sys.modules[spec.name] = mod
                spec.loader.exec_module(mod)
            except Exception:           # the reference file needs something that is absent here: leave its names out
                sys.modules.pop("model.render._reference_obj", None)
                mod = None
            _ref_cache[f] = mod
            return mod
    return None


def __getattr__(name):      # PEP 562
    if not name.startswith("__"):
        ref = _reference_obj()
        if ref is not None and hasattr(ref, name):
            return getattr(ref, name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))


def _host(t, dtype):
    """torch tensor / ndarray -> C-contiguous host ndarray of `dtype` (what `.detach().cpu().numpy()` yields, obj.py:135-141)."""
    if t is None:
        return None
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


def obj_text(v_pos, t_pos_idx, v_nrm=None, t_nrm_idx=None, v_tex=None, t_tex_idx=None, mtl_name="mesh", write_texcoords=True,
             threads=0):
    """The bytes `write_obj` puts in the .obj file for one instance: v_pos [V,3], v_tex [Vt,2], v_nrm [Vn,3] float32 and
    [F,3] integer index arrays (0-based).  `write_texcoords=False` = the reference's `save_material=False`: no 'vt' lines, the
    faces keep their texcoord column (obj.py:145,166)."""
    v_pos = _host(v_pos, np.float32)
    v_tex = _host(v_tex, np.float32)
    v_nrm = _host(v_nrm, np.float32)
    t_pos = _host(t_pos_idx, np.int64)
    t_tex = _host(t_tex_idx, np.int64) if v_tex is not None else None
    t_nrm = _host(t_nrm_idx, np.int64) if v_nrm is not None else None
    if v_pos.ndim != 2 or v_pos.shape[1] != 3 or t_pos.ndim != 2 or t_pos.shape[1] != 3:
        raise ValueError("obj_text: v_pos must be [V,3] and t_pos_idx [F,3]")
    if (v_tex is not None and (v_tex.ndim != 2 or v_tex.shape[1] != 2)) or (v_nrm is not None and (v_nrm.ndim != 2 or v_nrm.shape[1] != 3)):
        raise ValueError("obj_text: v_tex must be [Vt,2] and v_nrm [Vn,3]")
    if v_tex is not None and write_texcoords:
        assert len(t_pos) == len(t_tex)                 # obj.py:146
    if v_tex is not None and t_tex.shape != t_pos.shape:
        raise ValueError("obj_text: t_tex_idx must match t_pos_idx")
    if v_nrm is not None:
        assert len(t_pos) == len(t_nrm)                 # obj.py:151
        if t_nrm.shape != t_pos.shape:
            raise ValueError("obj_text: t_nrm_idx must match t_pos_idx")
    name = mtl_name.encode() if isinstance(mtl_name, str) else bytes(mtl_name)
    n_tex = len(v_tex) if (v_tex is not None and write_texcoords) else 0
    n_nrm = len(v_nrm) if v_nrm is not None else 0
    lib = _lib.lib()
    bound = C.c_size_t()
    _lib.check(lib.b2a_obj_text_bound(len(v_pos), n_tex, n_nrm, len(t_pos), len(name), C.byref(bound)))
    out = np.empty(bound.value, np.uint8)           # untouched pages beyond the text are never committed
    written = C.c_size_t()
    name_buf = C.create_string_buffer(name, len(name) + 1)
    ptr = lambda a: None if a is None else a.ctypes.data
    _lib.check(lib.b2a_obj_format(ptr(v_pos), len(v_pos), ptr(v_tex), n_tex, ptr(v_nrm), n_nrm, ptr(t_pos), ptr(t_tex), ptr(t_nrm),
                                  len(t_pos), C.cast(name_buf, C.c_void_p), len(name), out.ctypes.data, out.size, C.byref(written),
                                  int(threads)))
    return out[:written.value]


def write_obj(folder, fname, mesh, idx, save_material=True, feat=None, resolution=[256, 256]):
    """Same signature, file names, console lines and file bytes as the reference's write_obj (obj.py:128-177)."""
    obj_file = os.path.join(folder, fname + '.obj')
    print("Writing mesh: ", obj_file)
    v_pos = mesh.v_pos[idx] if mesh.v_pos is not None else None
    v_nrm = mesh.v_nrm[idx] if mesh.v_nrm is not None else None
    v_tex = mesh.v_tex[idx] if mesh.v_tex is not None else None
    t_pos_idx = mesh.t_pos_idx[0] if mesh.t_pos_idx is not None else None
    t_nrm_idx = mesh.t_nrm_idx[0] if mesh.t_nrm_idx is not None else None
    t_tex_idx = mesh.t_tex_idx[0] if mesh.t_tex_idx is not None else None
    text = obj_text(v_pos, t_pos_idx, v_nrm, t_nrm_idx, v_tex, t_tex_idx, mtl_name=fname, write_texcoords=bool(save_material))
    print("    writing %d vertices" % len(v_pos))
    if v_tex is not None and save_material:
        print("    writing %d texcoords" % len(v_tex))
    if v_nrm is not None:
        print("    writing %d normals" % len(v_nrm))
    print("    writing %d faces" % len(t_pos_idx))
    with open(obj_file, "wb") as f:
        f.write(memoryview(text))

    if save_material and mesh.material is not None:
        mtl_file = os.path.join(folder, fname + '.mtl')
        print("Writing material: ", mtl_file)
        material = importlib.import_module("model.render.material")     # the reference's own module (overlay mode)
        material.save_mtl(mtl_file, mesh.material, mesh=mesh.get_n(idx), feat=feat, resolution=resolution)

    print("Done exporting mesh")
