import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fm = importlib.import_module("3danimals_b200.field_mlp")
nets = importlib.import_module("3danimals_b200.networks")
cuda = torch.device("cuda:0")
torch.manual_seed(0)
mm = torch.tensor([[0., 1.]] * 9, device=cuda)
tex = nets.CoordMLP(3, 9, 8, nf=256, activation="sigmoid", min_max=mm, n_harmonic_functions=10, embedder_scalar=2 * np.pi / 7.0 * 0.9, extra_feat_dim=256, symmetrize=True).to(cuda)
dino = nets.CoordMLP(3, 16, 5, nf=256, activation="sigmoid", n_harmonic_functions=8, embedder_scalar=2 * np.pi / 7.0 * 0.9, symmetrize=True).to(cuda)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
x = (torch.rand(N, 3, device=cuda) - 0.5) * 6
img = torch.sort(torch.randint(0, 2, (N,), device=cuda)).values
feat = torch.randn(2, 256, device=cuda)
for net, f in ((tex, feat), (dino, None)):
    xr = x.clone().requires_grad_(True)
    out = fm.coord_mlp_rows(net, xr, f, img, 2)
    out.sum().backward()
    torch.cuda.synchronize()
    print("ok", tuple(out.shape), float(out.mean()), float(xr.grad.abs().max()))
