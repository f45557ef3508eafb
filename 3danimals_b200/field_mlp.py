"""The field MLPs of the pixel shader on the tensor cores: `CoordMLP.forward` (reference model/networks/MLPs.py:34-101) on
the covered rows as tcgen05 GEMMs of libb2a.so (csrc/field_mlp.cu), forward and backward.

`coord_mlp_rows(net, x, feat, img, n_img)` evaluates an UNMODIFIED `CoordMLP` module (the reference's class or this package's
twin: same parameters, read in place) on rows x [N,3]; a per-image feature enters the first hidden layer as a per-image bias
(the feature half of `Linear(nf + C -> nf)` runs on n_img rows in PyTorch, so its gradients flow through autograd).  Arithmetic:
every product is three bf16 MMAs over a hi / lo split of both fp32 operands with fp32 accumulation (passes = 3: ~5e-6 of the
result's magnitude, inside the fp32 configs' 1e-4 contract), or one (passes = 1) under autocast - the bird config wraps the fields
in fp16 autocast (train_magicpony_bird.yaml:52).  Pre-activations are kept in fp32 for the backward (ReLU masks, weight
gradients); the backward is one dgrad GEMM (ReLU-derivative mask in the epilogue) and one wgrad GEMM per layer.
"""
import os

import torch

from . import _lib
from . import ops

_call, _p, _stream = ops._call, ops._p, ops._stream

ops.KERNELS_PER_CALL.update({"b2a_mlp_pack_weights": 1, "b2a_mlp_rows_gemm": 1, "b2a_mlp_wgrad": 1, "b2a_mlp_embed_fwd": 1, "b2a_mlp_embed_bwd": 1,
                             "b2a_mlp_colsum_segments": 1})

# B2A_FIELD_MLP=torch keeps the fields on PyTorch's fp32 GEMMs (A/B measurements); default: the tcgen05 path
ENABLED = os.environ.get("B2A_FIELD_MLP", "tc") != "torch"


def _pack(w, n, k, transpose):
    """Shared-memory image (bf16 hi | lo, K-chunked) of B[n][k] = w[n, k] (or w[k, n] when transpose) for the GEMM's bulk copies."""
    nb = ops._size(ops._L().b2a_mlp_packed_bytes, n, k)
    buf = torch.empty(nb, dtype=torch.uint8, device=w.device)
    _call("b2a_mlp_pack_weights", (_p(w), w.stride(0), n, k, int(transpose), _p(buf), nb, _stream()))
    return buf


def _pack_many(jobs):
    """[(w, n, k, transpose), ...] -> the packed images, one allocation and ONE launch (b2a_mlp_pack_weights_many)."""
    import numpy as np
    sizes = [(ops._size(ops._L().b2a_mlp_packed_bytes, n, k) + 255) // 256 * 256 for _, n, k, _ in jobs]
    buf = torch.empty(sum(sizes), dtype=torch.uint8, device=jobs[0][0].device)
    base, off, views, table = buf.data_ptr(), 0, [], np.empty((len(jobs), 6), np.int64)
    for i, ((w, n, k, t), sz) in enumerate(zip(jobs, sizes)):
        table[i] = (w.data_ptr(), w.stride(0), n, k, int(t), base + off)
        views.append(buf[off:off + sz])
        off += sz
    _call("b2a_mlp_pack_weights_many", (table.ctypes.data, len(jobs), _stream()))
    return views


def _gemm(a, k, packed, n, relu, passes, epi, bias=None, bias_rows=None, mask=None, out=None, ldo=None, mask_bits=None, want_bits=False):
    """-> out, or (out, sign bits of out [rows, ceil16(n)/32 words... one uint32 per 32 columns]) with want_bits."""
    rows = a.shape[0]
    if out is None:
        out = torch.empty(rows, n if ldo is None else ldo, device=a.device)
    bits = torch.empty(rows, ((n + 15) // 16 * 16 + 31) // 32, dtype=torch.int32, device=a.device) if want_bits else None
    _call("b2a_mlp_rows_gemm", (_p(a), a.stride(0), rows, k, _p(packed), n, int(relu), passes, epi, _p(bias), _p(bias_rows), _p(mask),
                                0 if mask is None else mask.stride(0), _p(mask_bits), _p(bits), _p(out), out.stride(0), _stream()))
    return (out, bits) if want_bits else out


def _wgrad(p, relu_p, q, relu_q, m, n, passes, out, transpose_out=False):
    _call("b2a_mlp_wgrad", (_p(p), p.stride(0), int(relu_p), _p(q), q.stride(0), int(relu_q), p.shape[0], m, n, passes, _p(out), out.stride(0),
                            int(transpose_out), _stream()))


def supported(net, x, feat):
    """A `CoordMLP` this path can evaluate: fp32 CUDA parameters, harmonic embedding, hidden width <= 256 (a multiple of 32), no
    dropout, sigmoid or no output activation, at least two layers."""
    if not ENABLED or type(net).__name__ != "CoordMLP" or not x.is_cuda or x.dtype != torch.float32:
        return False
    layers = getattr(getattr(net, "mlp", None), "network", None)
    if layers is None or not hasattr(net, "in_layer") or getattr(net, "embedder", None) is None:
        return False
    lin = [m for m in layers if isinstance(m, torch.nn.Linear)]
    other = [m for m in layers if not isinstance(m, (torch.nn.Linear, torch.nn.ReLU, torch.nn.Sigmoid))]
    nf = net.in_layer.out_features
    if other or len(lin) < 2 or nf > 256 or nf % 32 or any(m.bias is not None for m in lin) or lin[-1].out_features > 256:
        return False
    if any(m.in_features != nf or m.out_features != nf for m in lin[1:-1]) or lin[-1].in_features != nf or lin[0].out_features != nf:
        return False
    extra = getattr(net, "extra_feat_dim", 0)
    if lin[0].in_features != nf + extra or (extra > 0) != (feat is not None):
        return False
    if net.in_layer.weight.dtype != torch.float32 or net.in_layer.in_features > 128:
        return False
    return True


class _FieldMLP(torch.autograd.Function):
    """rows x [N,3] -> out [N,cout].  Inputs after cfg: bias_img [n_img,nf] | None, w_in, b_in, w0 (full [nf, nf+extra]), hidden..., w_out."""

    @staticmethod
    def forward(ctx, x, img, seg_start, cfg, bias_img, w_in, b_in, *ws):
        n_harm, scalar, symmetrize, concat, passes, sigmoid = cfg
        x = ops._f32(x, "x")
        N = x.shape[0]
        nf, kin = w_in.shape
        dev = x.device
        st = _stream()
        ldE = (kin + 31) // 32 * 32
        E = torch.empty(N, ldE, device=dev)
        _call("b2a_mlp_embed_fwd", (_p(x), x.stride(0), N, n_harm, float(scalar), int(symmetrize), int(concat), _p(E), ldE, st))
        # every pre-activation is kept in fp32 (the weight gradients need relu(z)); its SIGN BITS are written by the same epilogue so
        # that the backward's ReLU-derivative mask reads 32 bytes per row instead of 1 KB
        cout = ws[-1].shape[0]
        packs = _pack_many([(w_in, nf, kin, False)] + [(w, nf, nf, False) for w in ws[:-1]] + [(ws[-1], cout, nf, False)])
        z, b = _gemm(E, kin, packs[0], nf, False, passes, 0, bias=b_in, want_bits=True)
        zs, bits = [z], [b]
        z, b = _gemm(zs[-1], nf, packs[1], nf, True, passes, 0, bias=bias_img, bias_rows=img if bias_img is not None else None,
                     want_bits=True)
        zs.append(z); bits.append(b)
        for i in range(1, len(ws) - 1):
            z, b = _gemm(zs[-1], nf, packs[1 + i], nf, True, passes, 0, want_bits=True)
            zs.append(z); bits.append(b)
        out = _gemm(zs[-1], nf, packs[-1], cout, True, passes, 2 if sigmoid else 0)
        ctx.save_for_backward(x, E, out, seg_start, w_in, *ws, *zs, *bits)
        ctx.cfg = cfg
        ctx.n_w = len(ws)
        ctx.has_bias = bias_img is not None
        ctx.n_img = 0 if bias_img is None else bias_img.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        n_harm, scalar, symmetrize, concat, passes, sigmoid = ctx.cfg
        saved = ctx.saved_tensors
        x, E, out, seg_start, w_in = saved[:5]
        ws = saved[5:5 + ctx.n_w]
        zs = saved[5 + ctx.n_w:5 + 2 * ctx.n_w]
        bits = saved[5 + 2 * ctx.n_w:]
        N = x.shape[0]
        nf, kin = w_in.shape
        cout = ws[-1].shape[0]
        dev = x.device
        st = _stream()
        need = ctx.needs_input_grad
        # output layer: dz = g * sigmoid', zero-padded to a multiple of 4 columns (row stride of the GEMM's A operand)
        cpad = (cout + 3) // 4 * 4
        dz = torch.zeros(N, cpad, device=dev)
        gz = g.float()
        dz[:, :cout] = gz * out * (1.0 - out) if sigmoid else gz
        # every weight gradient is accumulated with reductions: ONE zero fill for all of them (16-byte aligned slices of one buffer)
        sizes = [(w.numel() + 3) // 4 * 4 for w in ws] + [(w_in.numel() + 3) // 4 * 4]
        pool = torch.zeros(sum(sizes), device=dev)
        offs = [0]
        for n_ in sizes:
            offs.append(offs[-1] + n_)
        zeros_like = lambda i, w: pool[offs[i]:offs[i] + w.numel()].view(w.shape)
        # transposed weight images of every layer for the dgrad GEMMs: one launch (tp[i] for ws[i], tp[-1] for w_in)
        tp = _pack_many([(ws[i], nf, nf, True) for i in range(ctx.n_w - 1)] + [(ws[-1], nf, cout, True)] + ([(w_in, kin, nf, True)] if need[0] else []))
        d_ws = [None] * ctx.n_w
        d_ws[-1] = zeros_like(ctx.n_w - 1, ws[-1])
        _wgrad(zs[-1], True, dz, False, nf, cout, passes, d_ws[-1], transpose_out=True)          # d_W_out[c, j] = sum_r dz[r, c] relu(z)[r, j]
        dzl = _gemm(dz, cout, tp[ctx.n_w - 1], nf, False, passes, 1, mask_bits=bits[-1])   # (dz . W_out) * [z_last > 0]
        for i in range(ctx.n_w - 2, 0, -1):                                                    # hidden layers W_i: z_{i+1} = W_i relu(z_i)
            d_ws[i] = zeros_like(i, ws[i])
            _wgrad(dzl, False, zs[i], True, nf, nf, passes, d_ws[i])
            dzl = _gemm(dzl, nf, tp[i], nf, False, passes, 1, mask_bits=bits[i])
        # first hidden layer: only its h half [nf, :nf] is a GEMM here; the feature half is the per-image bias (PyTorch side)
        d_ws[0] = zeros_like(0, ws[0])
        _wgrad(dzl, False, zs[0], True, nf, nf, passes, d_ws[0])
        d_bias = None
        if ctx.has_bias and need[4]:
            d_bias = torch.empty(ctx.n_img, nf, device=dev)
            _call("b2a_mlp_colsum_segments", (_p(dzl), dzl.stride(0), _p(seg_start), ctx.n_img, nf, _p(d_bias), st))
        dz0 = _gemm(dzl, nf, tp[0], nf, False, passes, 1, mask_bits=bits[0])
        d_w_in = zeros_like(ctx.n_w, w_in)
        _wgrad(dz0, False, E, False, nf, kin, passes, d_w_in)
        d_b_in = torch.empty(1, nf, device=dev)                  # column sums of dz0 = one segment [0, N)
        whole = torch.tensor([0, N], dtype=torch.int64, device=dev) if seg_start is None else torch.stack([seg_start[0], seg_start[-1]])
        _call("b2a_mlp_colsum_segments", (_p(dz0), dz0.stride(0), _p(whole), 1, nf, _p(d_b_in), st))
        d_b_in = d_b_in.view(nf)
        d_x = None
        if need[0]:
            dE = _gemm(dz0, nf, tp[ctx.n_w], kin, False, passes, 0, ldo=E.shape[1])
            d_x = torch.empty(N, 3, device=dev)
            _call("b2a_mlp_embed_bwd", (_p(x), x.stride(0), N, n_harm, float(scalar), int(symmetrize), int(concat), _p(dE), dE.stride(0), _p(d_x), 3, st))
        return (d_x, None, None, None, d_bias, d_w_in, d_b_in) + tuple(d_ws)


def coord_mlp_rows(net, x, feat, img, n_img):
    """CoordMLP.forward (MLPs.py:72-98) on rows x [N,3] grouped by image (img [N] ascending; feat [n_img, C] or None)."""
    if not x.is_cuda:
        raise _lib.B2AError("field MLP rows must be CUDA tensors (the B200 hot path has no CPU fallback)")
    layers = net.mlp.network
    lin = [m for m in layers if isinstance(m, torch.nn.Linear)]
    sigmoid = any(isinstance(m, torch.nn.Sigmoid) for m in layers)
    nf = net.in_layer.out_features
    passes = 1 if torch.is_autocast_enabled() else 3
    emb = net.embedder
    freqs = emb.frequencies
    n_harm = int(freqs.numel())
    scalar = float(freqs[0]) if n_harm else 1.0
    bias_img, img32, seg = None, None, None
    if feat is not None:
        w0 = lin[0].weight
        bias_img = torch.nn.functional.linear(torch.relu(feat.float()), w0[:, nf:])          # [n_img, nf]: the feature half, with autograd
        img32 = img.to(torch.int32)
        # first row of every image (img is ascending): a binary search per image instead of a histogram of all the rows
        seg = torch.searchsorted(img, torch.arange(n_img + 1, dtype=img.dtype, device=x.device))
    cfg = (n_harm, scalar, bool(net.symmetrize), bool(net.embed_concat_pts), passes, sigmoid)
    out = _FieldMLP.apply(x, img32, seg, cfg, bias_img, net.in_layer.weight, net.in_layer.bias, *[m.weight for m in lin])
    if net.min_max is not None:
        out = out * (net.min_max[:, 1] - net.min_max[:, 0]) + net.min_max[:, 0]
    return out
