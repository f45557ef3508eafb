"""Articulation-angle constraints on the host: the numpy oracle and the product's stage tables against the reference's own methods
(tests/golden/articulation.npz: InstancePredictorBase.apply_articulation_constraints :435-511 and Fauna's sequence, run on CPU by
tests/golden/make_goldens.py).  The only arithmetic that differs between libraries is tanh itself (numpy / glibc vs torch / Sleef:
one ulp); with the reference's tanh both restatements are BIT-exact, which is what pins the stage-table factorisation."""
import importlib.util
import os

import numpy as np
import torch

from conftest import GOLDEN, pkg
from oracle import articulation_ref as A


def _configs():
    spec = importlib.util.spec_from_file_location("make_goldens_cfg", os.path.join(GOLDEN, "make_goldens.py"))
    src = open(spec.origin).read()
    start, stop = src.index("def articulation_configs"), src.index("def articulation_case")
    ns = {}
    exec(src[start:stop], ns)
    return ns["articulation_configs"]()


def _torch_tanh(a):
    return torch.tanh(torch.from_numpy(np.ascontiguousarray(a))).numpy()


def _stages_on_host(st, x, g):
    """What csrc/articulation.cu does, in numpy fp32 (each stage individually rounded), with the reference's tanh."""
    v = x.copy()
    for t in st.pre:
        v = v * t
    t_ = _torch_tanh(v)
    y = t_.copy()
    for f, div in zip(st.post, st.div):
        y = y / f if div else y * f
    d = g.copy()
    for f, div in list(zip(st.post, st.div))[::-1]:
        d = d / f if div else d * f
    d = d * (1.0 - t_.astype(np.float64) ** 2).astype(np.float32)     # 1 - t*t with one rounding: torch's tanh_backward fuses it
    for t in st.pre[::-1]:
        d = d * t
    return y, d


def test_oracle_matches_reference_methods():
    d = np.load(os.path.join(GOLDEN, "articulation.npz"))
    for name, cfg, add, it in _configs():
        x, want = d[name + ":x"], d[name + ":y"]
        run = (lambda **kw: A.base(x, cfg, **kw)) if add is None else (lambda **kw: A.fauna(x, cfg, add, it, **kw))
        assert np.array_equal(run(tanh=_torch_tanh), want), name          # bit-exact with the reference's tanh
        assert np.abs(run() - want).max() <= 2.4e-7, name                  # one ulp of tanh in (0.5, 1] times <= 1.05
        assert np.abs(want).max() > 0.1, name


def test_stage_tables_reproduce_reference_bits():
    P = pkg("predictors")
    d = np.load(os.path.join(GOLDEN, "articulation.npz"))
    for name, cfg, add, it in _configs():
        K = cfg.num_body_bones + cfg.num_leg_bones * cfg.num_legs
        st = P.base_stages(cfg, K) if add is None else P.fauna_stages(cfg, add, K, it)
        assert len(st.post) == len(st.div) and len(st.post) <= 24 and all(t.shape == (K, 3) and t.dtype == np.float32 for t in st.pre + st.post)
        y, dx = _stages_on_host(st, d[name + ":x"], d[name + ":g"])
        assert np.array_equal(y, d[name + ":y"]), name
        # the gradient too: autograd's chain is the same stages backwards (its tanh' is grad * (1 - y*y), fp32)
        assert np.array_equal(dx, d[name + ":d_x"]), name


def test_stage_tables_reject_missing_bones():
    P = pkg("predictors")
    from types import SimpleNamespace as NS
    cfg = NS(num_body_bones=8, num_leg_bones=3, num_legs=4, output_multiplier=0.1, static_root_bones=False, constrain_legs=True,
             use_fauna_constraints=True, extra_constraints=False, max_arti_angle=60)
    try:
        P.base_stages(cfg, 12)          # the reference's index assignment raises IndexError for bone 19 of 12
    except IndexError:
        return
    raise AssertionError("expected IndexError")
