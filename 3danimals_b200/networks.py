"""Minimal PyTorch coordinate-field MLP for the standalone bench / tests.

The field networks are OUT of the hot-path scope ("host code stays PyTorch"); when this package is
overlaid on the reference tree, DMTetGeometry uses the reference's own `model.networks.CoordMLP` unchanged.  This
module only provides a module of the same architecture and parameter names (reference model/networks/MLPs.py:34-101:
harmonic embedding -> Linear(in_layer) -> ReLU -> bias-free Linear stack) so the benchmark can run without the
reference tree (which is absent on the GPU box).
"""
import torch
import torch.nn as nn


class HarmonicEmbedding(nn.Module):
    """sin/cos of x * scalar * 2^k, k < n (reference model/networks/HarmonicEmbedding.py:8-46)."""

    def __init__(self, n_harmonic_functions=10, scalar=1.0):
        super().__init__()
        self.register_buffer("frequencies", scalar * (2.0 ** torch.arange(n_harmonic_functions)), persistent=False)

    def forward(self, x):
        e = (x[..., None] * self.frequencies).reshape(*x.shape[:-1], -1)
        return torch.cat((e.sin(), e.cos()), dim=-1)


class MLP(nn.Module):
    def __init__(self, cin, cout, num_layers, nf=256, dropout=0, activation=None):
        super().__init__()
        assert num_layers >= 1
        dims = [cin] + [nf] * (num_layers - 1) + [cout]
        layers = []
        for i in range(num_layers):
            if i > 0:
                layers.append(nn.ReLU(inplace=True))
            layers.append(nn.Linear(dims[i], dims[i + 1], bias=False))
            if dropout and 0 < i < num_layers - 1:
                layers.append(nn.Dropout(dropout))
        if activation == "sigmoid":
            layers.append(nn.Sigmoid())
        elif activation == "tanh":
            layers.append(nn.Tanh())
        elif activation == "relu":
            layers.append(nn.ReLU(inplace=True))
        elif activation is not None:
            raise NotImplementedError(activation)
        self.network = nn.Sequential(*layers)

    def forward(self, x):
        return self.network(x)


class CoordMLP(nn.Module):
    def __init__(self, cin, cout, num_layers, nf=256, dropout=0, activation=None, min_max=None, n_harmonic_functions=10,
                 embedder_scalar=1, embed_concat_pts=True, extra_feat_dim=0, symmetrize=False, in_layer_relu=False):
        super().__init__()
        self.extra_feat_dim = extra_feat_dim
        self.embed_concat_pts = embed_concat_pts
        if n_harmonic_functions and n_harmonic_functions > 0:
            self.embedder = HarmonicEmbedding(n_harmonic_functions, embedder_scalar)
            dim_in = cin * 2 * n_harmonic_functions + (cin if embed_concat_pts else 0)
        else:
            self.embedder = None
            dim_in = cin
        self.in_layer = nn.Linear(dim_in, nf)
        self.mlp = MLP(nf + extra_feat_dim, cout, num_layers, nf, dropout, activation)
        self.symmetrize = symmetrize
        self.in_layer_relu = in_layer_relu
        self.bsdf = None
        if min_max is not None:
            self.register_buffer("min_max", min_max)
        else:
            self.min_max = None

    def forward(self, x, feat=None):
        if self.symmetrize:
            x = torch.cat([x[..., :1].abs(), x[..., 1:]], -1)
        h = x
        if self.embedder is not None:
            h = self.embedder(x)
            if self.embed_concat_pts:
                h = torch.cat([x, h], -1)
        h = self.in_layer(h)
        if self.in_layer_relu:
            h = torch.relu(h)
        if feat is not None:
            while feat.dim() < h.dim():
                feat = feat.unsqueeze(1)
            h = torch.cat([h, feat.expand(*h.shape[:-1], -1)], -1)
        out = self.mlp(torch.relu(h))
        if self.min_max is not None:
            out = out * (self.min_max[:, 1] - self.min_max[:, 0]) + self.min_max[:, 0]
        return out

    def sample(self, x, feat=None):
        return self.forward(x, feat)


class CoordMLP_Mod(nn.Module):
    """The weight-modulated variant (reference MLPs.py:104-247, Fauna) is only available from the reference tree."""

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("CoordMLP_Mod comes from the reference's model.networks (overlay mode)")
