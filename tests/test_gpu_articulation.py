"""b2a_articulation_constraints_fwd/bwd through 3danimals_b200.predictors against the reference's own methods (golden fixture from
InstancePredictorBase.apply_articulation_constraints :435-511 and Fauna's sequence) and the numpy oracle.  Bar: the stage chain is
bit-exact given the same tanh (tests/test_oracle_articulation.py); CUDA's tanhf and the reference's CPU tanh differ by <= 2 ulp, so
values <= 3e-7 absolute (|out| <= 1.05) and gradients <= 2e-6 of the largest gradient."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from conftest import GOLDEN, pkg
from oracle import articulation_ref as A
from test_oracle_articulation import _configs

pytestmark = pytest.mark.gpu


def _run(P, cfg, add, it, x, g):
    self = NS(cfg_articulation=cfg, cfg_additional=add)
    xt = x.clone().requires_grad_(True)
    y = P.apply_articulation_constraints(self, xt) if add is None else P.fauna_articulation_angles(self, xt, it)
    y.backward(g)
    return self, y.detach(), xt.grad


def test_matches_reference_methods(cuda):
    P, ops = pkg("predictors"), pkg("ops")
    d = np.load(os.path.join(GOLDEN, "articulation.npz"))
    for name, cfg, add, it in _configs():
        x, g = torch.from_numpy(d[name + ":x"]).to(cuda), torch.from_numpy(d[name + ":g"]).to(cuda)
        ops.stats.reset()
        self, y, dx = _run(P, cfg, add, it, x, g)
        assert ops.stats.calls == {"b2a_articulation_constraints_fwd": 1, "b2a_articulation_constraints_bwd": 1}, name
        assert float((y.cpu() - torch.from_numpy(d[name + ":y"])).abs().max()) <= 3e-7, name
        want = d[name + ":d_x"]
        assert float(np.abs(dx.cpu().numpy() - want).max()) <= 2e-6 * np.abs(want).max(), name
        assert np.array_equal(y.cpu().numpy() == 0, d[name + ":y"] == 0), name            # masked entries are exact zeros
        if add is not None:
            assert self.constrain_legs == (it <= add.iter_leg_rotation_start)


def test_large_batch_against_oracle_and_method_install(cuda):
    P = pkg("predictors")
    name, cfg, add, it = _configs()[0]
    rng = np.random.RandomState(0)
    x = (rng.randn(64, 10, 20, 3) * 12).astype(np.float32)
    x[0, 0, :, :] = 0
    x[1, 0, 0, :] = [1e4, -1e4, 1e-30]                # saturated tanh, denormal-scale input
    want = A.base(x, cfg)

    class Predictor:                                  # stands for InstancePredictorBase: the method reads cfg_articulation only
        cfg_articulation = cfg

        def apply_articulation_constraints(self, articulation_angles, **kwargs):
            raise AssertionError("not replaced")

    P.install(Predictor)
    xt = torch.from_numpy(x).to(cuda)
    before = xt.clone()
    y = Predictor().apply_articulation_constraints(xt, total_iter=5)
    assert tuple(y.shape) == x.shape and torch.equal(xt, before)             # out of place
    assert float(np.abs(y.cpu().numpy() - want).max()) <= 3e-7
    # the tables are cached per (config, bone count, device)
    n = len(P._TABLES)
    Predictor().apply_articulation_constraints(xt)
    assert len(P._TABLES) == n


def test_errors(cuda):
    P, lib = pkg("predictors"), pkg("_lib")
    _, cfg, _, _ = _configs()[0]
    self = NS(cfg_articulation=cfg)
    with pytest.raises(ValueError):
        P.apply_articulation_constraints(self, torch.zeros(2, 20, 3, device=cuda))
    with pytest.raises(IndexError):
        P.apply_articulation_constraints(self, torch.zeros(2, 1, 12, 3, device=cuda))         # 20 bones configured, 12 given
    with pytest.raises((lib.B2AError, ValueError, TypeError, RuntimeError)):
        P.apply_articulation_constraints(self, torch.zeros(2, 1, 20, 3))                       # CPU tensor: no fallback
