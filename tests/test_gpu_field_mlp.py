"""The tcgen05 field MLP (csrc/field_mlp.cu through 3danimals_b200.field_mlp) against the reference's own `CoordMLP`
(model/networks/MLPs.py:34-101; the package's state-dict-compatible twin when the reference tree is absent) evaluated in fp64.

Bar: outputs <= 1e-4 of the output range at default initialisation AND with the hidden weights scaled x2.24 (a stand-in for a
trained network: pre-activations grow, which is where single-pass bf16 / tf32 operands fail, profiles/mlp_precision_study_r1.txt);
gradients in the relative L2 norm per tensor (a ReLU pre-activation within rounding of zero takes the other branch in any fp32
evaluation - torch's own fp32 forward included - and moves single entries of the earlier layers' gradients)."""
import copy

import numpy as np
import pytest
import torch

from conftest import pkg

pytestmark = pytest.mark.gpu


def _nets(cuda, scale):
    nets = pkg("networks")
    torch.manual_seed(0)
    mm = torch.tensor([[0., 1.]] * 9, device=cuda)
    tex = nets.CoordMLP(3, 9, 8, nf=256, activation="sigmoid", min_max=mm, n_harmonic_functions=10, embedder_scalar=2 * np.pi / 7.0 * 0.9,
                        extra_feat_dim=256, symmetrize=True).to(cuda)
    dino = nets.CoordMLP(3, 16, 5, nf=256, activation="sigmoid", n_harmonic_functions=8, embedder_scalar=2 * np.pi / 7.0 * 0.9, symmetrize=True).to(cuda)
    small = nets.CoordMLP(3, 1, 3, nf=64, activation=None, n_harmonic_functions=4, embedder_scalar=1.0, symmetrize=False, extra_feat_dim=32).to(cuda)
    with torch.no_grad():
        for net in (tex, dino, small):
            for m in net.mlp.network:
                if isinstance(m, torch.nn.Linear):
                    m.weight.mul_(scale)
    return tex, dino, small


@pytest.mark.parametrize("scale", [1.0, 2.24])
def test_field_mlp_matches_fp64_reference(cuda, scale):
    fm = pkg("field_mlp")
    B = 3
    for net, feat_dim in zip(_nets(cuda, scale), (256, 0, 32)):
        torch.manual_seed(1)
        N = 5000
        x = (torch.rand(N, 3, device=cuda) - 0.5) * 6
        x[:50] = 0                                            # rows shaded at gb_tex_pos = 0 exist under msaa
        img = torch.sort(torch.randint(0, B, (N,), device=cuda)).values
        feat = torch.randn(B, feat_dim, device=cuda) if feat_dim else None
        assert fm.supported(net, x, feat)
        xr = x.clone().requires_grad_(True)
        fr = feat.clone().requires_grad_(True) if feat is not None else None
        net.zero_grad()
        out = fm.coord_mlp_rows(net, xr, fr, img, B)
        g = torch.randn_like(out)
        out.backward(g)
        got = [p.grad.double().clone() for p in net.parameters()] + [xr.grad.double()] + ([fr.grad.double()] if fr is not None else [])
        n64 = copy.deepcopy(net).double()
        n64.zero_grad()
        x64 = x.double().requires_grad_(True)
        f64 = feat.double().requires_grad_(True) if feat is not None else None
        ref = n64.sample(x64, feat=None if f64 is None else f64.index_select(0, img))
        ref.backward(g.double())
        want = [p.grad for p in n64.parameters()] + [x64.grad] + ([f64.grad] if f64 is not None else [])
        rng = float(ref.abs().max())
        err = float((out.double() - ref).abs().max()) / rng
        assert err < 1e-4, (type(net).__name__, scale, err)
        names = [n for n, _ in net.named_parameters()] + ["x"] + (["feat"] if fr is not None else [])
        errs = {n: float((a - b).norm() / b.norm().clamp_min(1e-30)) for n, a, b in zip(names, got, want)}
        last = [n for n in names if n.startswith("mlp.network")][-1]
        # the output layer's weight gradient involves no ReLU decision downstream of a rounding difference: forward-grade accuracy;
        # every earlier tensor also sees the handful of ReLU branches that flip at this (5e-6) operand precision
        assert errs[last] < 2e-4, (scale, errs)
        assert max(errs.values()) < 1e-2, (scale, errs)


def test_field_mlp_single_pass_under_autocast(cuda):
    """passes = 1 (one bf16 MMA per product) is what runs under autocast - the bird config's arithmetic; bf16-grade agreement."""
    fm = pkg("field_mlp")
    tex, dino, _ = _nets(cuda, 1.0)
    x = (torch.rand(2000, 3, device=cuda) - 0.5) * 6
    img = torch.zeros(2000, dtype=torch.long, device=cuda)
    ref = dino.sample(x)
    with torch.autocast("cuda", dtype=torch.float16):
        out = fm.coord_mlp_rows(dino, x, None, img, 1)
    assert out.dtype == torch.float32 and float((out - ref).abs().max()) < 2e-2


def test_render_uses_the_tensor_core_fields(cuda):
    """render_mesh routes CoordMLP fields through the tcgen05 path on covered rows; images and gradients against the PyTorch fields."""
    pipe, fm, ops = pkg("pipeline"), pkg("field_mlp"), pkg("ops")
    torch.manual_seed(0)
    sc = pipe.SyntheticScene(grid_res=32, batch=2, image_res=64, sdf_noise=0.0)
    hp = pipe.HotPath(sc, cuda, mlps=True)
    g1, g2 = sc.upstream_grads()
    d1, d2 = torch.from_numpy(g1).to(cuda) * 1e3, torch.from_numpy(g2).to(cuda) * 1e3
    res = {}
    for on in (True, False):
        fm.ENABLED = on
        try:
            hp.zero_grad()
            ops.stats.reset()
            d_sdf, d_ang = hp.step(d1, d2)
            calls = dict(ops.stats.calls)
            shaded, dino = hp.forward()
            res[on] = (shaded.detach().clone(), dino.detach().clone(), d_sdf.clone(), d_ang.clone(), calls)
        finally:
            fm.ENABLED = True
    assert res[True][4].get("b2a_mlp_rows_gemm", 0) > 20 and res[False][4].get("b2a_mlp_rows_gemm", 0) == 0
    for a, b in zip(res[True][:2], res[False][:2]):
        assert float((a - b).abs().max()) < 1e-4 * max(float(b.abs().max()), 1e-12)
    for a, b in zip(res[True][2:4], res[False][2:4]):
        assert float((a - b).norm() / b.norm()) < 2e-3
