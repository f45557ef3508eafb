"""Articulation-angle constraints in front of skinning, as one kernel per direction.

Reference: `InstancePredictorBase.apply_articulation_constraints` (model/predictors/InstancePredictorBase.py:435-511) and Fauna's
three-part form (`InstancePredictorFauna.forward_articulation` :225-234 with `apply_articulation_constraints` :187-212 and
`apply_fauna_articulation_regularizer` :149-185).  Both are "scale, tanh, then a config-dependent chain of per-(bone, axis) factors";
the reference writes the chain as ~40 mask-algebra tensor statements per call.  Here the config is turned ONCE into stage tables
`[stage, bone, axis]` (`base_stages` / `fauna_stages`, cached per config and bone count) and `b2a_articulation_constraints_fwd/bwd`
apply them in the reference's order, each stage one individually rounded fp32 operation - the same arithmetic, one launch.

Use from the unmodified reference tree (callers are not overlaid - this is one assignment a maintainer adds):

    from model.predictors.InstancePredictorBase import InstancePredictorBase
    import importlib; importlib.import_module("3danimals_b200.predictors").install(InstancePredictorBase)

For Fauna the four statements `*= output_multiplier; tanh(); apply_articulation_constraints(); apply_fauna_articulation_regularizer()`
(:228-234) become `articulation_angles = predictors.fauna_articulation_angles(self, articulation_angles, total_iter)`.
There is no fallback: CPU tensors raise (ops._f32), as everywhere in this package.
"""
import numpy as np
import torch

from . import ops

_TABLES = {}


class Stages:
    """pre / post: float32 arrays [n, K, 3]; div: which post stages divide."""

    def __init__(self, K):
        self.K, self.pre, self.post, self.div = int(K), [], [], []

    def _table(self, fill=1.0):
        return np.full((self.K, 3), fill, np.float32)

    def all(self, where, factor, div=False):
        getattr(self, where).append(self._table(factor))
        if where == "post":
            self.div.append(bool(div))

    def some(self, where, bones, axes, factor, others=1.0):
        """`factor` at (bones x axes), `others` elsewhere; bones outside [0, K) raise like the reference's index assignment does."""
        t = self._table(others)
        bones = [int(b) for b in bones]
        if bones and (max(bones) >= self.K or min(bones) < -self.K):
            raise IndexError("bone index out of range for %d bones" % self.K)
        for a in axes:
            t[bones, a] = factor
        getattr(self, where).append(t)
        if where == "post":
            self.div.append(False)
        return t

    def finish(self, device):
        pre = torch.from_numpy(np.stack(self.pre) if self.pre else np.zeros((0, self.K, 3), np.float32)).to(device)
        post = torch.from_numpy(np.stack(self.post) if self.post else np.zeros((0, self.K, 3), np.float32)).to(device)
        mask = sum(1 << j for j, d in enumerate(self.div) if d)
        return pre, post, mask


def _angle_scale(st, max_arti_angle):
    # `articulation_angles * max_arti_angle / 180 * np.pi`: three fp32 operations, left to right
    st.all("post", float(max_arti_angle))
    st.all("post", 180.0, div=True)
    st.all("post", float(np.pi))


def base_stages(cfg, K):
    """InstancePredictorBase.py:435-511 as stages.  cfg: the predictor's `cfg_articulation`."""
    st = Stages(K)
    nb = int(cfg.num_body_bones)
    st.all("pre", float(cfg.output_multiplier))                                   # :436
    if cfg.static_root_bones:                                                     # :437-441 (a 0/1 mask, all three axes)
        st.some("pre", [nb // 2 - 1, nb - 1], (0, 1, 2), 0.0)
    if cfg.constrain_legs:                                                        # :443-453
        legs = nb + np.arange(int(cfg.num_leg_bones) * int(cfg.num_legs))
        st.some("post", legs, (2,), 0.3)
        st.some("post", legs, (1,), 0.3)
        if cfg.use_fauna_constraints:                                             # :457-484
            top = [10, 13, 16, 19]
            st.some("post", top, (1, 2), 0.05)
            st.some("post", top, (0,), 0.75)
            t = st.some("post", [8, 9, 11, 12, 14, 15, 17, 18], (1, 2), 0.0)
            t[[8, 9, 11, 12, 14, 15, 17, 18], 0] = 0.3
            st.some("post", list(range(8)), (2,), 0.1)
    if getattr(cfg, "extra_constraints", False):                                  # :486-508
        half = int(cfg.num_leg_bones) * int(cfg.num_legs) // 2
        legs = [nb + i for i in range(half)] + [nb + half + i for i in range(half)]
        st.some("post", legs, (2,), 0.3)
        st.some("post", legs, (1,), 0.3)
        st.some("post", [8, 11, 14, 17], (1, 2), 0.05)
        st.some("post", [9, 10, 12, 13, 15, 16, 18, 19], (1, 2), 0.0)
    _angle_scale(st, cfg.max_arti_angle)                                          # :510
    return st


def fauna_stages(cfg, cfg_additional, K, total_iter):
    """InstancePredictorFauna.py:228-229 (scale, tanh), :187-212 (constraints), :149-185 (regulariser) as stages."""
    st = Stages(K)
    nb = int(cfg.num_body_bones)
    st.all("pre", float(cfg.output_multiplier))
    if cfg.static_root_bones:                                                     # after the tanh here (:188-192)
        st.some("post", [nb // 2 - 1, nb - 1], (0, 1, 2), 0.0)
    start = cfg_additional.iter_leg_rotation_start
    if total_iter <= start:                                                       # :194-209
        legs = nb + np.arange(int(cfg.num_leg_bones) * int(cfg.num_legs))
        st.some("post", legs, (2,), 0.3)
        st.some("post", legs, (1,), 0.3)
    if start > 0 and total_iter > start and cfg_additional.forbid_leg_rotate:     # :152-168
        if cfg_additional.small_leg_angle:
            st.some("post", [8, 11, 14, 17], (1, 2), 0.05)
        st.some("post", [9, 10, 12, 13, 15, 16, 18, 19], (1, 2), 0.0)
    _angle_scale(st, cfg.max_arti_angle)                                          # :171
    mult = cfg_additional.reg_body_rotate_mult * 180 * 1.0 / (cfg.max_arti_angle * np.pi)   # :176-177 (host double)
    st.some("post", list(range(8)), (2,), float(mult))                            # :179-182
    return st


class _Constraints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pre, post, div_mask):
        x = ops._f32(x, "articulation_angles")
        K = x.shape[-2]
        rows = x.numel() // max(K * 3, 1)
        out = torch.empty_like(x)
        ops._call("b2a_articulation_constraints_fwd", (ops._p(x), ops._p(pre) if pre.numel() else None, pre.shape[0],
                                                       ops._p(post) if post.numel() else None, post.shape[0], div_mask, rows, K, ops._p(out),
                                                       ops._stream()))
        ctx.save_for_backward(x, pre, post)
        ctx.div_mask = div_mask
        return out

    @staticmethod
    def backward(ctx, g):
        x, pre, post = ctx.saved_tensors
        K = x.shape[-2]
        rows = x.numel() // max(K * 3, 1)
        g = ops._f32(g, "d_articulation_angles")
        d_x = torch.empty_like(x)
        ops._call("b2a_articulation_constraints_bwd", (ops._p(x), ops._p(pre) if pre.numel() else None, pre.shape[0],
                                                       ops._p(post) if post.numel() else None, post.shape[0], ctx.div_mask, rows, K, ops._p(g),
                                                       ops._p(d_x), ops._stream()))
        return d_x, None, None, None


def _cfg_key(cfg, names):
    return tuple((n, getattr(cfg, n, None)) for n in names)


_BASE_KEYS = ("num_body_bones", "num_leg_bones", "num_legs", "output_multiplier", "static_root_bones", "constrain_legs", "use_fauna_constraints",
              "extra_constraints", "max_arti_angle")


def _apply(x, key, build):
    if x.dim() != 4 or x.shape[-1] != 3:
        raise ValueError("articulation_angles must be [B, F, K, 3], got %s" % (tuple(x.shape),))
    key = key + (x.shape[2], str(x.device))
    tabs = _TABLES.get(key)
    if tabs is None:
        if len(_TABLES) > 256:
            _TABLES.clear()
        tabs = _TABLES[key] = build(x.shape[2]).finish(x.device)
    return _Constraints.apply(x, *tabs)


def apply_articulation_constraints(self, articulation_angles, **kwargs):
    """Drop-in for the METHOD InstancePredictorBase.apply_articulation_constraints (:435-511); `self` needs `cfg_articulation` only.
    Out of place (the reference scales its argument in place first; no caller reads the argument afterwards, :527, :542)."""
    cfg = self.cfg_articulation
    return _apply(articulation_angles, ("base",) + _cfg_key(cfg, _BASE_KEYS), lambda K: base_stages(cfg, K))


def fauna_articulation_angles(self, articulation_angles, total_iter):
    """InstancePredictorFauna.forward_articulation's four statements between the articulation network and skinning (:228-234),
    including the `self.constrain_legs` flag the reference method leaves behind (:194-197)."""
    cfg, add = self.cfg_articulation, self.cfg_additional
    self.constrain_legs = bool(total_iter <= add.iter_leg_rotation_start)
    phase = (bool(total_iter <= add.iter_leg_rotation_start), bool(add.iter_leg_rotation_start > 0 and total_iter > add.iter_leg_rotation_start))
    key = ("fauna", phase) + _cfg_key(cfg, _BASE_KEYS) + _cfg_key(add, ("forbid_leg_rotate", "small_leg_angle", "reg_body_rotate_mult"))
    return _apply(articulation_angles, key, lambda K: fauna_stages(cfg, add, K, total_iter))


def install(predictor_cls):
    """Replace `apply_articulation_constraints` on the reference's InstancePredictorBase (or a subclass that inherits it)."""
    predictor_cls.apply_articulation_constraints = apply_articulation_constraints
    return predictor_cls
