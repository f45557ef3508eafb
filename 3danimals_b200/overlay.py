"""Module substitution: make the reference tree import this package's hot path (SURVEY.md §8b "what calls it").

    import importlib; importlib.import_module("3danimals_b200.overlay").install()

After `install()`, `import model.geometry.dmtet`, `model.geometry.skinning`, `model.render.mesh`,
`model.render.render`, `model.render.light`, `model.render.obj` and `nvdiffrast(.torch)` resolve to the B200-native modules, so the reference's
`model/models`, `model/predictors`, Trainer and visualization scripts run unmodified on top of libb2a.so.
Everything else under `model.*` (networks, light, material, util, ...) keeps loading from the reference tree.
"""
import importlib
import importlib.abc
import importlib.util
import sys

_PKG = __name__.rsplit(".", 1)[0]

ALIASES = {
    "model.geometry.dmtet": _PKG + ".geometry.dmtet",
    "model.geometry.skinning": _PKG + ".geometry.skinning",
    "model.render.mesh": _PKG + ".render.mesh",
    "model.render.render": _PKG + ".render.render",
    "model.render.light": _PKG + ".render.light",
    "model.render.obj": _PKG + ".render.obj",
    "nvdiffrast": _PKG + ".nvdiffrast_shim",
    "nvdiffrast.torch": _PKG + ".nvdiffrast_shim.torch",
}


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname in ALIASES:
            is_pkg = fullname == "nvdiffrast"
            return importlib.util.spec_from_loader(fullname, _AliasLoader(ALIASES[fullname]), is_package=is_pkg)
        return None


_finder = None


def install():
    """Idempotent.  Must run before the reference's `model` package is imported."""
    global _finder
    if _finder is None:
        _finder = _AliasFinder()
        sys.meta_path.insert(0, _finder)
    for name in ALIASES:
        if name in sys.modules and getattr(sys.modules[name], "__name__", "") != ALIASES[name]:
            del sys.modules[name]
    return _finder


def uninstall():
    global _finder
    if _finder is not None:
        sys.meta_path.remove(_finder)
        _finder = None
    for name, target in ALIASES.items():
        if name in sys.modules and getattr(sys.modules[name], "__name__", "") == target:
            del sys.modules[name]
