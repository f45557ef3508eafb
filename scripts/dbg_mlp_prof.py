import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("3danimals_b200._lib"); lib = L.lib()
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
rows = 210000
A = torch.randn(rows, 256, device=dev); W = torch.randn(256, 256, device=dev) / 16
nb = ctypes.c_size_t(0); lib.b2a_mlp_packed_bytes(256, 256, ctypes.byref(nb))
Wp = torch.empty(nb.value, dtype=torch.uint8, device=dev)
lib.b2a_mlp_pack_weights(W.data_ptr(), 256, 256, 256, 0, Wp.data_ptr(), Wp.numel(), st)
out = torch.empty(rows, 256, device=dev)
for _ in range(3):
    lib.b2a_mlp_rows_gemm(A.data_ptr(), 256, rows, 256, Wp.data_ptr(), 256, 1, 3, 0, None, None, None, 0, None, None, out.data_ptr(), 256, st)
torch.cuda.synchronize()
P2 = torch.randn(rows, 256, device=dev); o2 = torch.zeros(256, 256, device=dev)
for _ in range(2):
    lib.b2a_mlp_wgrad(P2.data_ptr(), 256, 0, A.data_ptr(), 256, 1, rows, 256, 256, 3, o2.data_ptr(), 256, 0, st)
torch.cuda.synchronize()
