"""Import the reference's own caller modules - UNMODIFIED - on top of the B200 drop-in (TEST INFRASTRUCTURE ONLY).

`load()` installs `3danimals_b200.overlay` and imports, from the staged tree (`oracle/stage_ref.py`; `/root/reference` in the
build container), the files that CALL the hot path:

    model/predictors/BasePredictorBase.py      BasePredictorBase.forward -> DMTetGeometry.getMesh          (R1)
    model/predictors/InstancePredictorBase.py  forward_articulation / get_bones / apply_articulation_constraints (R4, R5, R3)
    model/models/AnimalModel.py                AnimalModel.render                                            (R6-R10)
    model/models/Fauna.py                      FaunaModel.get_random_view_mask (second ['shaded'] render)

Only what this image cannot provide is stubbed, and only at import level: the third-party packages the reference imports but
the image lacks (pytorch3d, matplotlib, imageio, omegaconf, accelerate), and `model/__init__.py` (it imports the Trainer ->
accelerate/hydra) is skipped by registering `model`, `model.models`, `model.predictors` as bare packages whose `__path__`
points into the tree.  Two reference files use a dataclass INSTANCE as a dataclass default (`InstancePredictorFauna.py:26`,
`InstancePredictorMotionVAE.py:24`), which Python >= 3.11 rejects; while those files are imported `dataclasses.dataclass` is
wrapped so that such a default becomes a `default_factory` (the value the reference's Python 3.10 would have used).  No byte of
a reference file is changed.  The classes are instantiated WITHOUT their constructors where the constructor needs the network
(`ViTEncoder` -> torch.hub): tests attach the attributes the called methods read.
"""
import copy
import dataclasses
import importlib
import os
import sys
import types

from . import stage_ref

_STUBS = ["pytorch3d", "pytorch3d.transforms", "matplotlib", "matplotlib.pyplot", "imageio", "omegaconf", "omegaconf.errors", "accelerate",
          "lpips", "trimesh", "xatlas", "moviepy", "moviepy.editor", "configargparse", "hydra"]
_BARE = {"model": "model", "model.models": "model/models", "model.predictors": "model/predictors"}
_state = None


def available():
    return stage_ref.available()


def _compat_dataclass(real):
    def wrap(cls=None, **kw):
        def fix(c):
            for name in list(getattr(c, "__annotations__", {})):
                v = c.__dict__.get(name)
                if dataclasses.is_dataclass(v) and not isinstance(v, type):
                    setattr(c, name, dataclasses.field(default_factory=lambda v=v: copy.deepcopy(v)))
            return real(c, **kw)
        return fix if cls is None else fix(cls)
    return wrap


def load():
    """-> namespace(root, BPB, IPB, AM, Fauna, networks, misc, render_util): the reference's modules bound to the overlay."""
    global _state
    if _state is not None:
        return _state
    if not available():
        raise RuntimeError("no reference tree: run `python oracle/stage_ref.py` in the build container")
    root = stage_ref.root()
    ov = importlib.import_module("3danimals_b200.overlay")
    saved = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.") or k.split(".")[0] in {s.split(".")[0] for s in _STUBS}}
    ov.install()
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = types.ModuleType(name)
                m.__path__ = []
                sys.modules[name] = m
                if "." in name:
                    setattr(sys.modules[name.rsplit(".", 1)[0]], name.rsplit(".", 1)[1], m)
    sys.modules["omegaconf.errors"].ConfigAttributeError = getattr(sys.modules["omegaconf.errors"], "ConfigAttributeError", AttributeError)
    for name, rel in _BARE.items():
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(root, rel)]
        sys.modules[name] = pkg
        if "." in name:
            setattr(sys.modules["model"], name.split(".")[1], pkg)
    ns = types.SimpleNamespace(root=root, saved=saved)
    ns.networks = importlib.import_module("model.networks")
    ns.misc = importlib.import_module("model.utils.misc")
    ns.render_util = importlib.import_module("model.render.util")
    preds = sys.modules["model.predictors"]
    real_dc = dataclasses.dataclass
    try:
        dataclasses.dataclass = _compat_dataclass(real_dc)
        for leaf in ("BasePredictorBase", "BasePredictorBank", "InstancePredictorBase", "InstancePredictorMotionVAE", "InstancePredictorFauna"):
            try:
                mod = importlib.import_module("model.predictors." + leaf)
            except Exception as e:          # MotionVAE / Fauna predictors are optional for the tests that use this harness
                setattr(ns, leaf + "_error", repr(e))
                continue
            for k, v in vars(mod).items():   # what `from .X import *` in model/predictors/__init__.py exports
                if not k.startswith("_"):
                    setattr(preds, k, v)
            setattr(ns, leaf, mod)
        ns.BPB, ns.IPB = ns.BasePredictorBase, ns.InstancePredictorBase
        ns.AM = importlib.import_module("model.models.AnimalModel")
        try:
            ns.Fauna = importlib.import_module("model.models.Fauna")
        except Exception as e:
            ns.Fauna, ns.Fauna_error = None, repr(e)
    finally:
        dataclasses.dataclass = real_dc
    _state = ns
    return ns


def unload():
    """Remove the reference modules and the overlay from sys.modules (tests call this in a finally block)."""
    global _state
    if _state is None:
        return
    importlib.import_module("3danimals_b200.overlay").uninstall()
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        del sys.modules[k]
    for name in _STUBS:
        m = sys.modules.get(name)
        if m is not None and getattr(m, "__file__", None) is None and getattr(m, "__path__", None) == []:
            del sys.modules[name]
    sys.modules.update(_state.saved)
    _state = None


def bare(cls):
    """An instance of a reference class whose constructor is NOT run (it would download a ViT): nn.Module state only (the model
    classes - AnimalModel, FaunaModel - are plain Python classes, the predictors are nn.Modules)."""
    import torch
    obj = cls.__new__(cls)
    if isinstance(obj, torch.nn.Module):
        torch.nn.Module.__init__(obj)
    return obj
