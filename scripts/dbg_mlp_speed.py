"""Field-MLP GEMM kernels against a device copy of the same bytes (event-timed, operands larger than L2)."""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("3danimals_b200._lib"); lib = L.lib()
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 184705
A = torch.randn(rows, 256, device=dev); W = torch.randn(256, 256, device=dev) / 16
nb = ctypes.c_size_t(0); lib.b2a_mlp_packed_bytes(256, 256, ctypes.byref(nb))
Wp = torch.empty(nb.value, dtype=torch.uint8, device=dev)
lib.b2a_mlp_pack_weights(W.data_ptr(), 256, 256, 256, 0, Wp.data_ptr(), Wp.numel(), st)
out = torch.empty(rows, 256, device=dev)
bits = torch.empty(rows, 8, dtype=torch.int32, device=dev)
P2 = torch.randn(rows, 256, device=dev); o2 = torch.zeros(256, 256, device=dev)

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

lib.b2a_mlp_rows_gemm(A.data_ptr(), 256, rows, 256, Wp.data_ptr(), 256, 1, 3, 0, None, None, None, 0, None, None, out.data_ptr(), 256, st)
ref = torch.relu(A).double() @ W.double().t()
print("gemm check (3-pass, relu on load): max |err| / max |ref| = %.2e   (B2A_MLP_CLUSTER=%s)" % (float((out.double() - ref).abs().max() / ref.abs().max()), os.environ.get("B2A_MLP_CLUSTER", "-")))
del ref
mb = rows * 256 * 4 / 1e6
print("rows %d: one [rows,256] fp32 matrix = %.1f MB" % (rows, mb))
def rep(name, us, nbytes): print("%-46s %8.1f us   %6.0f GB/s" % (name, us, nbytes / us / 1e3))
rep("torch copy_ (read + write)", timed(lambda: out.copy_(A)), 2 * mb * 1e6)
rep("torch relu out= (read + write)", timed(lambda: torch.clamp_min(A, 0.0, out=out)), 2 * mb * 1e6)
rep("torch sum over rows (read)", timed(lambda: A.sum(0)), mb * 1e6)
for passes in (3, 1):
    rep("gemm fwd relu_on_load passes=%d (read + write)" % passes, timed(lambda: lib.b2a_mlp_rows_gemm(A.data_ptr(), 256, rows, 256, Wp.data_ptr(), 256, 1, passes, 0, None, None, None, 0, None, None, out.data_ptr(), 256, st)), 2 * mb * 1e6)
    rep("gemm fwd + sign bits out passes=%d" % passes, timed(lambda: lib.b2a_mlp_rows_gemm(A.data_ptr(), 256, rows, 256, Wp.data_ptr(), 256, 1, passes, 0, None, None, None, 0, None, bits.data_ptr(), out.data_ptr(), 256, st)), 2 * mb * 1e6)
    rep("gemm dgrad mask_bits passes=%d" % passes, timed(lambda: lib.b2a_mlp_rows_gemm(A.data_ptr(), 256, rows, 256, Wp.data_ptr(), 256, 0, passes, 1, None, None, None, 0, bits.data_ptr(), None, out.data_ptr(), 256, st)), 2 * mb * 1e6)
    rep("wgrad passes=%d (2 reads)" % passes, timed(lambda: lib.b2a_mlp_wgrad(P2.data_ptr(), 256, 0, A.data_ptr(), 256, 1, rows, 256, 256, passes, o2.data_ptr(), 256, 0, st)), 2 * mb * 1e6)
seg16 = torch.linspace(0, rows, 17, device=dev).long(); o16 = torch.empty(16, 256, device=dev)
rep("colsum_segments 16 segments (read)", timed(lambda: lib.b2a_mlp_colsum_segments(A.data_ptr(), 256, seg16.data_ptr(), 16, 256, o16.data_ptr(), st)), mb * 1e6)
seg1 = torch.tensor([0, rows], device=dev); o1 = torch.empty(1, 256, device=dev)
rep("colsum_segments 1 segment (read)", timed(lambda: lib.b2a_mlp_colsum_segments(A.data_ptr(), 256, seg1.data_ptr(), 1, 256, o1.data_ptr(), st)), mb * 1e6)
want = A.double().sum(0)
print("colsum check: max rel err %.2e / %.2e" % (float((o1[0].double() - want).abs().max() / want.abs().max()), float((o16.double().sum(0) - want).abs().max() / want.abs().max())))
torch.backends.cuda.matmul.allow_tf32 = True
rep("cuBLAS tf32 A @ W (read + write)", timed(lambda: torch.mm(A, W, out=out)), 2 * mb * 1e6)
Ah, Wh = A.bfloat16(), W.bfloat16(); oh = torch.empty(rows, 256, device=dev, dtype=torch.bfloat16)
rep("cuBLAS bf16 A @ W (bf16 read + write)", timed(lambda: torch.mm(Ah, Wh, out=oh)), mb * 1e6)
