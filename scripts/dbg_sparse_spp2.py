import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pipe = importlib.import_module("3danimals_b200.pipeline")
R = importlib.import_module("3danimals_b200.render.render")
cuda = torch.device("cuda:0")
torch.manual_seed(0)
sc = pipe.SyntheticScene(grid_res=32, batch=2, image_res=64, sdf_noise=0.0)
hp = pipe.HotPath(sc, cuda, mlps=True)
hp.spp = 2
# capture what _sample_field sees
calls = []
orig = R._sample_field
def spy(net, gb, feat, sparse):
    out = orig(net, gb, feat, sparse)
    if net is hp.material:
        out.retain_grad()
        calls.append((gb.detach().clone(), sparse, out))
    return out
R._sample_field = spy
g1, g2 = sc.upstream_grads()
d1, d2 = torch.from_numpy(g1).to(cuda) * 1e3, torch.from_numpy(g2).to(cuda) * 1e3
outs = {}
for sparse in (True, False):
    hp.sparse_fields = sparse
    hp.zero_grad()
    calls.clear()
    hp.step(d1, d2)
    gb, sp, out = calls[0]
    outs[sparse] = (gb, sp, out.detach().clone(), out.grad.clone())
gbS, spS, oS, gS = outs[True]
gbD, _, oD, gD = outs[False]
print("gb equal", torch.equal(gbS, gbD))
idx = spS[0]
sel = torch.zeros(gbS.shape[0] * gbS.shape[1] * gbS.shape[2], dtype=torch.bool, device=cuda); sel[idx] = True
gDf = gD.reshape(-1, gD.shape[-1]); gSf = gS.reshape(-1, gS.shape[-1])
print("rows selected", int(sel.sum()), "dense rows with nonzero grad", int((gDf.abs().sum(1) > 0).sum()), "nonzero outside sel", int(((gDf.abs().sum(1) > 0) & ~sel).sum()))
print("grad diff on sel rows", float((gDf[sel] - gSf[sel]).abs().max()), "max", float(gDf.abs().max()))
print("out diff on sel rows", float((oD.reshape(-1, 9)[sel] - oS.reshape(-1, 9)[sel]).abs().max()))
zero_rows = (gbS.reshape(-1, 3)[sel].abs().sum(1) == 0)
print("selected rows with x == 0:", int(zero_rows.sum()))
bad = (gDf[sel] - gSf[sel]).abs().sum(1) > 1e-9
print("rows with differing grad:", int(bad.sum()), "of which x==0:", int((bad & zero_rows).sum()))
