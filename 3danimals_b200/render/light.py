"""Drop-in for the reference's model/render/light.py: `DirectionalLight` with the per-pixel shading arithmetic
(light.py:186-193: ambient + diffuse * clamp(dot(dir, n), 0), times kd) as ONE sm_100a kernel per direction
(ops.shade_directional -> csrc/shade.cu) instead of ~5 elementwise torch kernels forward and ~10 backward over
[B,H,W,3].  The light-direction MLP (feat -> 4 sigmoid outputs -> upper-hemisphere direction + two intensities,
light.py:177-184) stays a PyTorch module with the reference's parameter names, so checkpoints load unchanged.

Every other name of the reference module (EnvironmentLight, load_env, create_trainable_env_rnd, ...: used by no
shipped config) is re-exported from the reference's own file when the reference tree is importable
(overlay mode); standalone, only DirectionalLight exists.
"""
import importlib.util
import os
import sys

import torch
import torch.nn.functional as F

from .. import ops


_ref_cache = {}


def _reference_light():
    """The reference's own model/render/light.py, loaded under a private name the first time one of its other names is asked
    for (overlay mode: `model.render` is then the reference's package); None when the reference tree is not importable."""
    pkg = sys.modules.get("model.render")
    for d in getattr(pkg, "__path__", None) or []:
        f = os.path.join(d, "light.py")
        if f in _ref_cache:
            return _ref_cache[f]
        if os.path.isfile(f):
            try:
                spec = importlib.util.spec_from_file_location("model.render._reference_light", f)
                mod = importlib.util.module_from_spec(spec)
                sys.modules[spec.name] = mod
                spec.loader.exec_module(mod)
            except Exception:           # the reference file needs something that is absent here: leave its names out
                sys.modules.pop("model.render._reference_light", None)
                mod = None
            _ref_cache[f] = mod
            return mod
    return None


def __getattr__(name):      # PEP 562: every name this module does not define resolves to the reference module's
    if not name.startswith("__"):
        ref = _reference_light()
        if ref is not None and hasattr(ref, name):
            return getattr(ref, name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))


def _mlp_class():
    nets = sys.modules.get("model.networks")
    if nets is not None and hasattr(nets, "MLP"):
        return nets.MLP                  # overlay mode: the reference's own MLP
    from ..networks import MLP
    return MLP


class DirectionalLight(torch.nn.Module):
    """Same constructor, parameters (`mlp.*`, buffer `intensity_min_max`), `forward` and `shade` contract as the
    reference class (light.py:168-193)."""

    def __init__(self, mlp_in, mlp_layers, mlp_hidden_size, intensity_min_max=None):
        super().__init__()
        self.mlp = _mlp_class()(mlp_in, 4, mlp_layers, nf=mlp_hidden_size, activation="sigmoid")
        if intensity_min_max is not None:
            self.register_buffer("intensity_min_max", intensity_min_max)
        else:
            self.intensity_min_max = None

    def forward(self, feat):
        out = self.mlp(feat)
        # direction in the upper hemisphere: (2a-1, 0.5, 2b-1) normalised; intensities rescaled into [min, max]
        d = torch.stack([out[..., 0] * 2 - 1, torch.full_like(out[..., 0], 0.5), out[..., 1] * 2 - 1], dim=-1)
        light_dir = F.normalize(d, dim=-1)
        inten = out[..., 2:]
        if self.intensity_min_max is not None:
            lo, hi = self.intensity_min_max[:, 0], self.intensity_min_max[:, 1]
            inten = inten * (hi - lo) + lo
        self.light_params = torch.cat([light_dir, inten], -1)
        return self.light_params

    def shade(self, feat, kd, normal):
        light_params = self.forward(feat)
        return ops.shade_directional(kd, normal, light_params.reshape(-1, 5))
