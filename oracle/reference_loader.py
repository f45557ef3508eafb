"""Load the reference's own geometry / shading files BY PATH (build container only).

Test infrastructure only (see oracle/__init__.py).  ``/root/reference`` does not exist
on the GPU box, so this module is used exclusively by ``oracle/make_goldens.py`` and by
CPU tests that skip when the reference tree is absent.

Recipe: stub the third-party modules the reference imports but this
image lacks, register bare namespace packages so ``model/__init__.py`` (Trainer, accelerate,
py3.12-incompatible dataclasses) never executes, then exec the individual files.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B2A_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "geometry", "dmtet.py")):
    # GPU box: the byte-identical staged copy (oracle/stage_ref.py -> git-ignored oracle/_ref, travels with the snapshot)
    _staged = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    if os.path.isfile(os.path.join(_staged, "model", "geometry", "dmtet.py")):
        REFERENCE_ROOT = _staged

_STUBS = [
    "nvdiffrast", "nvdiffrast.torch", "imageio", "matplotlib", "matplotlib.pyplot",
    "pytorch3d", "accelerate", "omegaconf", "omegaconf.errors",
]
_PKGS = {
    "model": "model",
    "model.geometry": "model/geometry",
    "model.render": "model/render",
    "model.networks": "model/networks",
    "model.utils": "model/utils",
    "model.render.renderutils": "model/render/renderutils",
}
_cache = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "geometry", "dmtet.py"))


def _load(name, relpath):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns a namespace with the reference modules: dmtet, skinning, geo_util, mesh, bsdf, rutil, mlps, light."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    saved = {k: sys.modules.get(k) for k in list(_STUBS) + list(_PKGS)}
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["omegaconf.errors"].ConfigAttributeError = AttributeError
    for name, rel in _PKGS.items():
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
        sys.modules[name] = pkg
    ns = types.SimpleNamespace()
    try:
        _load("model.networks.util", "model/networks/util.py")
        _load("model.networks.HarmonicEmbedding", "model/networks/HarmonicEmbedding.py")
        ns.mlps = _load("model.networks.MLPs", "model/networks/MLPs.py")
        for sym in ("CoordMLP", "CoordMLP_Mod", "MLP"):
            if hasattr(ns.mlps, sym):
                setattr(sys.modules["model.networks"], sym, getattr(ns.mlps, sym))
        ns.rutil = _load("model.render.util", "model/render/util.py")
        sys.modules["model.render.obj"] = types.ModuleType("model.render.obj")
        sys.modules["model.render"].obj = sys.modules["model.render.obj"]
        sys.modules["model.render"].util = ns.rutil
        ns.mesh = _load("model.render.mesh", "model/render/mesh.py")
        sys.modules["model.render"].mesh = ns.mesh
        ns.dmtet = _load("model.geometry.dmtet", "model/geometry/dmtet.py")
        ns.geo_util = _load("model.geometry.util", "model/geometry/util.py")
        sys.modules["model.geometry"].util = ns.geo_util
        ns.skinning = _load("model.geometry.skinning", "model/geometry/skinning.py")
        ns.bsdf = _load("model.render.renderutils.bsdf", "model/render/renderutils/bsdf.py")
        try:    # light.py only needs its imports to resolve (nvdiffrast / renderutils are stubs); DirectionalLight is pure torch
            ns.light = _load("model.render.light", "model/render/light.py")
        except Exception:
            ns.light = None
    finally:
        # leave the reference modules importable under their own names only inside `ns`;
        # restore sys.modules so the product overlay (3danimals_b200.overlay) is not shadowed.
        for k in list(sys.modules):
            if k == "model" or k.startswith("model."):
                del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            elif k in sys.modules and k in _STUBS:
                del sys.modules[k]
    _cache = ns
    return ns


def reference_dmtet(device="cpu"):
    """The reference DMTet configured for CPU (dmtet.py:23-46; map_uv reads self.device, :72-73)."""
    ns = load()
    mt = ns.dmtet.DMTet(device=device)
    mt.device = device
    return mt


def reference_write_obj():
    """The reference's own `write_obj` (model/render/obj.py:128-177), loaded by path.  Its sibling modules (texture, mesh,
    material: GPU / nvdiffrast / imageio dependent) are never touched when `mesh.material is None`, so empty stand-ins satisfy
    the file's imports.  sys.modules is restored afterwards."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    names = ["model", "model.render", "model.render.texture", "model.render.mesh", "model.render.material", "model.render.obj"]
    saved = {k: sys.modules.get(k) for k in names}
    try:
        for name in names[:-1]:
            mod = types.ModuleType(name)
            if name in ("model", "model.render"):
                mod.__path__ = [os.path.join(REFERENCE_ROOT, name.replace(".", "/"))]
            sys.modules[name] = mod
        for leaf in ("texture", "mesh", "material"):
            setattr(sys.modules["model.render"], leaf, sys.modules["model.render." + leaf])
        return _load("model.render.obj", "model/render/obj.py").write_obj
    finally:
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
