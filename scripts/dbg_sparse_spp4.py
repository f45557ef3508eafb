import importlib, sys, os, copy
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pipe = importlib.import_module("3danimals_b200.pipeline")
cuda = torch.device("cuda:0")
print("matmul precision", torch.get_float32_matmul_precision(), "allow_tf32", torch.backends.cuda.matmul.allow_tf32, os.environ.get("NVIDIA_TF32_OVERRIDE"))
torch.manual_seed(0)
sc = pipe.SyntheticScene(grid_res=32, batch=2, image_res=64, sdf_noise=0.0)
hp = pipe.HotPath(sc, cuda, mlps=True)
net = hp.material
N = 1536
def manual(net, x, feat, variant):
    xs = torch.cat([x[..., :1].abs(), x[..., 1:]], -1)
    h = torch.cat([xs, net.embedder(xs)], -1)
    h = net.in_layer(h)
    layers = net.mlp.network
    if variant == "concat":
        z = F.linear(torch.relu(torch.cat([h, feat], -1)), layers[0].weight)
    else:
        w = layers[0].weight
        z = F.linear(torch.relu(h), w[:, :256]) + F.linear(torch.relu(feat), w[:, 256:])
    for i in range(1, len(layers)):
        m = layers[i]
        z = torch.relu(z) if isinstance(m, torch.nn.ReLU) else m(z)
    return z * (net.min_max[:, 1] - net.min_max[:, 0]) + net.min_max[:, 0]
for nzero in (0, 117):
    x = torch.randn(N, 3, device=cuda)
    x[:nzero] = 0
    img = (torch.arange(N, device=cuda) >= N // 2).long()
    feat = hp.feat.index_select(0, img)
    g = torch.randn(N, 9, device=cuda)
    def grads(fn, net, dt):
        net.zero_grad()
        out = fn(net, x.to(dt), feat.to(dt))
        (out * g.to(dt)).sum().backward()
        return [(n, p.grad.double().clone()) for n, p in net.named_parameters()]
    n64 = copy.deepcopy(net).double()
    ref = grads(lambda n, x, f: n.sample(x, feat=f), n64, torch.float64)
    for name, fn in (("sample", lambda n, x, f: n.sample(x, feat=f)), ("concat", lambda n, x, f: manual(n, x, f, "concat")), ("split", lambda n, x, f: manual(n, x, f, "split"))):
        got = grads(fn, net, torch.float32)
        print(nzero, name, " ".join("%.1e" % ((a - c).abs().max() / c.abs().max()) for (_, a), (_, c) in zip(got, ref)))
