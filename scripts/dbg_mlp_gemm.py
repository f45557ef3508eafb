"""Standalone check of the tcgen05 rows GEMM (b2a_mlp_rows_gemm) against fp64 torch."""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("3danimals_b200._lib")
lib = L.lib()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream

def packed(W, transpose=False):
    N, K = (W.shape[1], W.shape[0]) if transpose else W.shape
    nb = ctypes.c_size_t(0)
    L.check(lib.b2a_mlp_packed_bytes(N, K, ctypes.byref(nb)))
    buf = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    L.check(lib.b2a_mlp_pack_weights(W.data_ptr(), W.stride(0), N, K, int(transpose), buf.data_ptr(), buf.numel(), st))
    return buf

def gemm(A, Wp, N, relu=False, passes=3, epi=0, bias=None, bias_rows=None, mask=None):
    rows, K = A.shape
    out = torch.full((rows, N), float("nan"), device=dev)
    L.check(lib.b2a_mlp_rows_gemm(A.data_ptr(), A.stride(0), rows, K, Wp.data_ptr(), N, int(relu), passes, epi,
                                  None if bias is None else bias.data_ptr(), None if bias_rows is None else bias_rows.data_ptr(),
                                  None if mask is None else mask.data_ptr(), 0 if mask is None else mask.stride(0), None, None, out.data_ptr(), out.stride(0), st))
    torch.cuda.synchronize()
    return out

torch.manual_seed(0)
for rows, K, N in ((128, 32, 16), (128, 32, 256), (300, 256, 256), (1000, 64, 256), (777, 256, 16), (210000, 256, 256)):
    A = torch.randn(rows, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    ref = A.double() @ W.double().t()
    for passes in (3, 1):
        out = gemm(A, packed(W), N, passes=passes)
        err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
        print("rows %6d K %3d N %3d passes %d  rel err %.2e  nan %d" % (rows, K, N, passes, err, int(torch.isnan(out).sum())))
# transposed weights, relu on load, per-image bias, mask, sigmoid
rows, K, N = 5000, 256, 256
A = torch.randn(rows, K, device=dev); W = torch.randn(K, N, device=dev) / 16
out = gemm(A, packed(W, transpose=True), N, relu=True)
ref = A.double().clamp_min(0) @ W.double()
print("transpose+relu  rel err %.2e" % ((out.double() - ref).abs().max() / ref.abs().max()).item())
W = torch.randn(N, K, device=dev) / 16
bias = torch.randn(7, N, device=dev); img = torch.randint(0, 7, (rows,), device=dev, dtype=torch.int32)
out = gemm(A, packed(W), N, bias=bias, bias_rows=img)
ref = A.double() @ W.double().t() + bias.double()[img.long()]
print("per-image bias  rel err %.2e" % ((out.double() - ref).abs().max() / ref.abs().max()).item())
mask = torch.randn(rows, N, device=dev)
out = gemm(A, packed(W), N, epi=1, mask=mask)
ref = (A.double() @ W.double().t()) * (mask > 0)
print("mask            rel err %.2e" % ((out.double() - ref).abs().max() / ref.abs().max()).item())
bv = torch.randn(9, device=dev); W9 = torch.randn(9, K, device=dev) / 16
out = gemm(A, packed(W9), 9, epi=2, bias=bv)
ref = torch.sigmoid(A.double() @ W9.double().t() + bv.double())
print("sigmoid N=9     rel err %.2e" % ((out.double() - ref).abs().max() / ref.abs().max()).item())
# timing at the workload size
rows = 210000
A = torch.randn(rows, 256, device=dev); W = torch.randn(256, 256, device=dev) / 16; Wp = packed(W)
out = torch.empty(rows, 256, device=dev)
for passes in (3, 1):
    for _ in range(3):
        lib.b2a_mlp_rows_gemm(A.data_ptr(), 256, rows, 256, Wp.data_ptr(), 256, 1, passes, 0, None, None, None, 0, None, None, out.data_ptr(), 256, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.b2a_mlp_rows_gemm(A.data_ptr(), 256, rows, 256, Wp.data_ptr(), 256, 1, passes, 0, None, None, None, 0, None, None, out.data_ptr(), 256, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("passes %d: %.1f us per layer of %d rows: %.1f TFLOP/s (fp32-equivalent), %.1f TFLOP/s of bf16 MMAs, %.0f GB/s" %
          (passes, ms * 1e3, rows, 2 * rows * 256 * 256 / ms / 1e9, passes * 2 * rows * 256 * 256 / ms / 1e9, 2 * rows * 1024 / ms / 1e6))
torch.backends.cuda.matmul.allow_tf32 = False
e0.record()
for _ in range(20):
    torch.mm(A, W.t(), out=out)
e1.record(); torch.cuda.synchronize()
print("torch fp32 mm: %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))

# ---- wgrad / embedding / column sums -------------------------------------------------------------------------------
def wgrad(P, Q, M, N, relu_p=False, relu_q=False, passes=3, transpose_out=False):
    out = torch.zeros((N, M) if transpose_out else (M, N), device=dev)
    L.check(lib.b2a_mlp_wgrad(P.data_ptr(), P.stride(0), int(relu_p), Q.data_ptr(), Q.stride(0), int(relu_q), P.shape[0], M, N, passes,
                              out.data_ptr(), out.stride(0), int(transpose_out), st))
    torch.cuda.synchronize()
    return out

for rows, M, N in ((64, 128, 32), (1000, 256, 256), (5000, 256, 64), (3333, 256, 16), (210000, 256, 256)):
    Pm = torch.randn(rows, M, device=dev); Qm = torch.randn(rows, N if N % 4 == 0 else N + 4 - N % 4, device=dev)
    ref = Pm.double().t() @ Qm.double()[:, :N].clamp_min(0)
    out = wgrad(Pm, Qm, M, N, relu_q=True)
    print("wgrad rows %6d M %3d N %3d rel err %.2e" % (rows, M, N, ((out.double() - ref).abs().max() / ref.abs().max()).item()))
rows = 4000
Pm = torch.randn(rows, 256, device=dev); Qm = torch.randn(rows, 16, device=dev)
out = wgrad(Pm, Qm, 256, 9, relu_p=True, transpose_out=True)
ref = (Pm.double().clamp_min(0).t() @ Qm.double()[:, :9]).t()
print("wgrad transposed out [9,256] rel err %.2e" % ((out.double() - ref).abs().max() / ref.abs().max()).item())
# timing
rows = 210000
Pm = torch.randn(rows, 256, device=dev); Qm = torch.randn(rows, 256, device=dev); o = torch.zeros(256, 256, device=dev)
for _ in range(3):
    lib.b2a_mlp_wgrad(Pm.data_ptr(), 256, 0, Qm.data_ptr(), 256, 1, rows, 256, 256, 3, o.data_ptr(), 256, 0, st)
e0.record()
for _ in range(20):
    lib.b2a_mlp_wgrad(Pm.data_ptr(), 256, 0, Qm.data_ptr(), 256, 1, rows, 256, 256, 3, o.data_ptr(), 256, 0, st)
e1.record(); torch.cuda.synchronize()
print("wgrad 256x256 over %d rows: %.1f us" % (rows, e0.elapsed_time(e1) / 20 * 1e3))
# embedding
x = torch.randn(5000, 3, device=dev) * 2
for nh, sym in ((10, 1), (8, 0)):
    ld = 64
    E = torch.empty(5000, ld, device=dev)
    sc = 2 * 3.14159265 / 7 * 0.9
    L.check(lib.b2a_mlp_embed_fwd(x.data_ptr(), 3, 5000, nh, ctypes.c_float(sc), sym, 1, E.data_ptr(), ld, st))
    xd = x.double().clone().requires_grad_(True)
    xs = torch.cat([xd[:, :1].abs(), xd[:, 1:]], -1) if sym else xd
    fr = sc * 2.0 ** torch.arange(nh, device=dev, dtype=torch.float64)
    em = (xs[..., None] * fr).view(5000, -1)
    ref = torch.cat([xs, em.sin(), em.cos()], -1)
    print("embed nh %d: err %.2e, pad zero %s" % (nh, (E[:, :ref.shape[1]].double() - ref).abs().max().item(), bool((E[:, ref.shape[1]:] == 0).all())))
    g = torch.randn(5000, ld, device=dev)
    dx = torch.empty(5000, 3, device=dev)
    L.check(lib.b2a_mlp_embed_bwd(x.data_ptr(), 3, 5000, nh, ctypes.c_float(sc), sym, 1, g.data_ptr(), ld, dx.data_ptr(), 3, st))
    (ref * g[:, :ref.shape[1]].double()).sum().backward()
    print("embed bwd: rel err %.2e" % ((dx.double() - xd.grad).abs().max() / xd.grad.abs().max()).item())
G = torch.randn(10000, 256, device=dev); seg = torch.tensor([0, 100, 100, 7000, 10000], device=dev, dtype=torch.int64)
o = torch.empty(4, 256, device=dev)
L.check(lib.b2a_mlp_colsum_segments(G.data_ptr(), 256, seg.data_ptr(), 4, 256, o.data_ptr(), st))
ref = torch.stack([G[a:b].double().sum(0) for a, b in ((0, 100), (100, 100), (100, 7000), (7000, 10000))])
print("colsum rel err %.2e" % ((o.double() - ref).abs().max() / ref.abs().max()).item())
