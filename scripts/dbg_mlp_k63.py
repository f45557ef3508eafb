import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("3danimals_b200._lib"); lib = L.lib()
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
rows, K, N, lda = int(sys.argv[1]), int(sys.argv[2]), 256, int(sys.argv[3])
A = torch.randn(rows, lda, device=dev); W = torch.randn(N, K, device=dev) / 8
bias = torch.randn(N, device=dev)
nb = ctypes.c_size_t(0); lib.b2a_mlp_packed_bytes(N, K, ctypes.byref(nb))
Wp = torch.empty(nb.value, dtype=torch.uint8, device=dev)
L.check(lib.b2a_mlp_pack_weights(W.data_ptr(), K, N, K, 0, Wp.data_ptr(), Wp.numel(), st))
out = torch.empty(rows, N, device=dev)
L.check(lib.b2a_mlp_rows_gemm(A.data_ptr(), lda, rows, K, Wp.data_ptr(), N, 0, 3, 0, bias.data_ptr(), None, None, 0, None, None, out.data_ptr(), N, st))
torch.cuda.synchronize()
ref = A[:, :K].double() @ W.double().t() + bias.double()
print("rel err %.2e" % ((out.double() - ref).abs().max() / ref.abs().max()).item())
