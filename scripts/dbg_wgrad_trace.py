"""clock64 phases of ONE CTA of mlp_wgrad_kernel (CTA 70, staging thread 0) at 185 k rows, 256 x 256.  Needs a trace build:
B2A_NVCC_DEFINES=-DB2A_MLP_TRACE python 3danimals_b200/build.py --force   (rebuild without it afterwards)."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("3danimals_b200._lib"); lib = L.lib()
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
rows = 184705
A = torch.randn(rows, 256, device=dev); P2 = torch.randn(rows, 256, device=dev); o2 = torch.zeros(256, 256, device=dev)
h = C.CDLL(L.LIB_PATH)
mhz = 1965.0
for it in range(3):
    lib.b2a_mlp_wgrad(P2.data_ptr(), 256, 0, A.data_ptr(), 256, 1, rows, 256, 256, 3, o2.data_ptr(), 256, 0, st)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 128)(); h.b2a_debug_mlp_trace(buf)
    t = np.array(buf[:], dtype=np.int64)
    us = lambda i, j: (t[i] - t[j]) / mhz
    print("run %d: chunk loop (39 chunks of 32 rows) %.2f us, wait for the last MMAs %.2f, epilogue (reductions into the shared result) %.2f, exit barrier %.2f"
          % (it, us(97, 96), us(98, 97), us(99, 98), us(100, 99)))
