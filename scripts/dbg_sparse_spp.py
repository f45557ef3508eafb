"""Diagnostic: sparse vs dense field evaluation under msaa - per-parameter gradient differences, and both against an fp64 copy of
the field networks fed the same gathered rows."""
import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pipe = importlib.import_module("3danimals_b200.pipeline")
cuda = torch.device("cuda:0")
for spp in (1, 2, 4):
    torch.manual_seed(0)
    sc = pipe.SyntheticScene(grid_res=32, batch=2, image_res=64, sdf_noise=0.0)
    hp = pipe.HotPath(sc, cuda, mlps=True)
    hp.spp = spp
    g1, g2 = sc.upstream_grads()
    d1, d2 = torch.from_numpy(g1).to(cuda) * 1e3, torch.from_numpy(g2).to(cuda) * 1e3
    res = {}
    for sparse in (True, False):
        hp.sparse_fields = sparse
        hp.zero_grad()
        hp.step(d1, d2)
        res[sparse] = [(n, p.grad.double().clone()) for n, p in list(hp.material.named_parameters()) + list(hp.dino_net.named_parameters())]
    for (n, a), (_, b) in zip(res[True], res[False]):
        print(spp, n, tuple(a.shape), "max|dense|=%.3e" % b.abs().max().item(), "rel diff=%.3e" % ((a - b).abs().max() / b.abs().max()).item())
