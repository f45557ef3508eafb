import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("3danimals_b200._lib"); lib = L.lib()
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
rows = 184705
A = torch.randn(rows, 256, device=dev); P2 = torch.randn(rows, 256, device=dev); o2 = torch.zeros(256, 256, device=dev)
h = C.CDLL(L.LIB_PATH)
for it in range(3):
    lib.b2a_mlp_wgrad(P2.data_ptr(), 256, 0, A.data_ptr(), 256, 1, rows, 256, 256, 3, o2.data_ptr(), 256, 0, st)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 128)(); h.b2a_debug_gemm_trace(buf)
    t = np.array(buf[:], dtype=np.int64); t0 = t[0]
    us = lambda i: (t[i] - t0) / 1965.0
    print("run %d: chunk loop %.2f us, final mma wait %.2f, epilogue (atomics) %.2f, exit barrier %.2f" % (it, us(1), us(2) - us(1), us(3) - us(2), us(4) - us(3)))
