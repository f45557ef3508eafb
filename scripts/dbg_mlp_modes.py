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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for passes, what in ((3, "full"), (1, "1 pass"), (0, "no MMA"), (-1, "no MMA, no A loads"), (-2, "3 pass, no output stores")):
    for _ in range(3):
        lib.b2a_mlp_rows_gemm(A.data_ptr(), 256, rows, 256, Wp.data_ptr(), 256, 1, passes, 0, None, None, None, 0, None, None, out.data_ptr(), 256, st)
    e0.record()
    for _ in range(20):
        lib.b2a_mlp_rows_gemm(A.data_ptr(), 256, rows, 256, Wp.data_ptr(), 256, 1, passes, 0, None, None, None, 0, None, None, out.data_ptr(), 256, st)
    e1.record(); torch.cuda.synchronize()
    print("%-28s %.1f us" % (what, e0.elapsed_time(e1) / 20 * 1e3))
# reference points: plain copy of the same bytes
e0.record()
for _ in range(20):
    out.copy_(A)
e1.record(); torch.cuda.synchronize()
print("torch copy 215 MB -> 215 MB   %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
