"""Host numerics study for SURVEY.md §8f-1 (tensor-core field MLP): which tcgen05 operand format keeps the texture field
(CoordMLP 8 x 256 -> 9, sigmoid, with a 256-d per-image feature; magicpony.yaml:65-74) inside the path's 1e-4 relative
tolerance?  The matmuls of every layer are emulated in numpy with the operand rounding of each candidate and fp32
accumulation (what TMEM accumulators do); everything else (harmonic embedding, ReLU, sigmoid) stays fp32.  Truth = fp64.

    python scripts/mlp_precision_study.py > profiles/mlp_precision_study_r1.txt

Formats: fp32 (cuBLAS SIMT today) | tf32 = 10 explicit mantissa bits (kind::tf32; operands truncated) | 3xtf32 = hi/lo split,
a_hi.b_hi + a_hi.b_lo + a_lo.b_hi | bf16 (kind::f16, 7 bits) | fp16 (kind::f16, 10 bits, the bird config's autocast) |
2xbf16 = bf16 hi/lo split of both operands, a_hi.b_hi + a_hi.b_lo + a_lo.b_hi (3 MMAs at the bf16 rate = 1.5 TF32 MMAs) |
3xbf16 = three-way bf16 split, the six terms a_i.b_j with i + j < 3.
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nets = importlib.import_module("3danimals_b200.networks")


def trunc_bits(x, keep):
    """fp32 -> keep `keep` explicit mantissa bits by truncation (tensor cores ignore the low operand bits)."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    return (u & np.uint32(0xFFFFFFFF << (23 - keep) & 0xFFFFFFFF)).view(np.float32)


def rne_bits(x, keep):
    """round-to-nearest-even to `keep` explicit mantissa bits (what a cvt.rn before the MMA does)."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    drop = 23 - keep
    u = u + ((1 << (drop - 1)) - 1) + ((u >> drop) & 1)
    return ((u >> drop) << drop).astype(np.uint32).view(np.float32)


def mm32(a, b):
    return (a.astype(np.float32) @ b.astype(np.float32)).astype(np.float32)


def split(x, keep, n):
    parts, r = [], x.astype(np.float32)
    for _ in range(n):
        h = rne_bits(r, keep)
        parts.append(h)
        r = (r - h).astype(np.float32)
    return parts


def matmul(a, b, mode):
    if mode == "fp64":
        return a.astype(np.float64) @ b.astype(np.float64)
    if mode == "fp32":
        return mm32(a, b)
    if mode == "tf32":
        return mm32(trunc_bits(a, 10), trunc_bits(b, 10))
    if mode == "tf32_rn":
        return mm32(rne_bits(a, 10), rne_bits(b, 10))
    if mode == "fp16":
        return mm32(a.astype(np.float16), b.astype(np.float16))
    if mode == "bf16":
        return mm32(rne_bits(a, 7), rne_bits(b, 7))
    if mode in ("3xtf32", "2xbf16", "3xbf16"):
        keep, n = (10, 2) if mode == "3xtf32" else (7, 2 if mode == "2xbf16" else 3)
        A, B = split(a, keep, n), split(b, keep, n)
        out = np.zeros((a.shape[0], b.shape[1]), np.float32)
        for i in range(n):                                  # terms a_i.b_j with i + j < n, smallest first
            for j in range(n):
                if i + j < n and not (i == 0 and j == 0):
                    out += mm32(A[i], B[j])
        return out + mm32(A[0], B[0])
    raise ValueError(mode)


def field_forward(net, x, feat, mode):
    f64 = mode == "fp64"
    dt = np.float64 if f64 else np.float32
    sd = {k: v.detach().numpy().astype(dt) for k, v in net.state_dict().items()}
    x = x.astype(dt)
    x = np.concatenate([np.abs(x[:, :1]), x[:, 1:]], -1) if net.symmetrize else x
    freqs = net.embedder.frequencies.numpy().astype(dt) if hasattr(net.embedder, "frequencies") else None
    e = (x[..., None] * freqs).reshape(x.shape[0], -1)
    h = np.concatenate([x, np.sin(e), np.cos(e)], -1)
    h = matmul(h, sd["in_layer.weight"].T, mode) + sd["in_layer.bias"]
    h = np.concatenate([h, feat.astype(dt)], -1)
    h = np.maximum(h, 0)
    keys = sorted((k for k in sd if k.startswith("mlp.network.") and k.endswith("weight")), key=lambda k: int(k.split(".")[2]))
    for i, k in enumerate(keys):
        h = matmul(h, sd[k].T, mode)
        if i + 1 < len(keys):
            h = np.maximum(h, 0)
    h = 1 / (1 + np.exp(-h))
    mm = sd["min_max"]
    return (h * (mm[:, 1] - mm[:, 0]) + mm[:, 0]).astype(dt)


def main():
    torch.manual_seed(0)
    mm = torch.tensor([[0.0, 1.0]] * 6 + [[-1.0, 1.0]] * 3)
    net = nets.CoordMLP(3, 9, 8, nf=256, activation="sigmoid", min_max=mm, n_harmonic_functions=10, embedder_scalar=2 * np.pi / 7 * 0.95,
                        embed_concat_pts=True, extra_feat_dim=256, symmetrize=True)
    rng = np.random.RandomState(0)
    N = 8192
    x = (rng.rand(N, 3).astype(np.float32) - 0.5) * 3.0
    feat = np.repeat(rng.randn(4, 256).astype(np.float32), N // 4, 0)
    # self-check of the emulation: the fp32 mode reproduces the module
    with torch.no_grad():
        ref = net(torch.from_numpy(x), feat=torch.from_numpy(feat)).numpy()
    truth = field_forward(net, x, feat, "fp64")
    print("texture field CoordMLP 8x256 -> 9 (+256-d feature), %d points, default torch init; truth = fp64" % N)
    print("self-check: |numpy fp32 emulation - torch module| max = %.2e" % np.abs(field_forward(net, x, feat, "fp32") - ref).max())
    print("%-9s %-14s %-14s %s" % ("operands", "max abs err", "max rel err*", "MMAs per product"))
    scale = np.abs(truth).max()
    for mode, mmas in (("fp32", "- (SIMT)"), ("3xtf32", "3"), ("3xbf16", "6"), ("2xbf16", "3"), ("tf32_rn", "1"), ("tf32", "1"), ("fp16", "1"), ("bf16", "1")):
        err = np.abs(field_forward(net, x, feat, mode).astype(np.float64) - truth).max()
        print("%-9s %-14.3e %-14.3e %s" % (mode, err, err / scale, mmas))
    print("* relative to the largest output magnitude (%.3f), the convention of the parity tests (conftest.rel_err)" % scale)
    # trained weights are larger than the default init: repeat with every hidden weight matrix scaled up (cumulative)
    total = 1.0
    for factor in (1.6, 1.4):
        total *= factor
        with torch.no_grad():
            for k, v in net.state_dict().items():
                if k.endswith("weight") and v.dim() == 2 and "mlp.network" in k:
                    v.mul_(factor)
        truth = field_forward(net, x, feat, "fp64")
        scale = np.abs(truth).max()
        print("\nsame network, every hidden weight matrix x%.2f (larger pre-activations, saturating sigmoids; output range %.3f..%.3f):"
              % (total, truth.min(), truth.max()))
        for mode in ("fp32", "3xtf32", "2xbf16", "tf32_rn", "fp16", "bf16"):
            err = np.abs(field_forward(net, x, feat, mode).astype(np.float64) - truth).max()
            print("%-9s %-14.3e %-14.3e" % (mode, err, err / scale))


if __name__ == "__main__":
    main()
