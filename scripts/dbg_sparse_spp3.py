import importlib, sys, os, copy
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pipe = importlib.import_module("3danimals_b200.pipeline")
R = importlib.import_module("3danimals_b200.render.render")
cuda = torch.device("cuda:0")
torch.manual_seed(0)
sc = pipe.SyntheticScene(grid_res=32, batch=2, image_res=64, sdf_noise=0.0)
hp = pipe.HotPath(sc, cuda, mlps=True)
net = hp.material
N = 1536
x = torch.randn(N, 3, device=cuda)
x[:117] = 0
img = (torch.arange(N, device=cuda) >= N // 2).long()
feat = hp.feat
g = torch.randn(N, 9, device=cuda)
def grads(fn, net):
    net.zero_grad()
    out = fn(net)
    (out * g.to(out.dtype)).sum().backward()
    return out.detach().double(), [(n, p.grad.double().clone()) for n, p in net.named_parameters()]
oa, ga = grads(lambda n: R._coord_mlp_rows(n, x, feat, img), net)
ob, gb = grads(lambda n: n.sample(x, feat=feat.index_select(0, img)), net)
net64 = copy.deepcopy(net).double()
oc, gc = grads(lambda n: n.sample(x.double(), feat=feat.double().index_select(0, img)), net64)
print("out: rows vs sample %.2e, rows vs f64 %.2e, sample vs f64 %.2e" % ((oa - ob).abs().max(), (oa - oc).abs().max(), (ob - oc).abs().max()))
for (n, a), (_, b), (_, c) in zip(ga, gb, gc):
    m = c.abs().max()
    print(n, "rows-sample %.2e rows-f64 %.2e sample-f64 %.2e" % ((a - b).abs().max() / m, (a - c).abs().max() / m, (b - c).abs().max() / m))
