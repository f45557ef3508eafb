"""clock64 timeline of ONE CTA of mlp_rows_gemm_kernel (CTA 700: staging thread 0 and the issuer thread) at 185 k rows, 256 -> 256.
Needs a trace build:  B2A_NVCC_DEFINES=-DB2A_MLP_TRACE python 3danimals_b200/build.py --force   (rebuild without it afterwards)."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("3danimals_b200._lib"); lib = L.lib()
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
rows = 184705
A = torch.randn(rows, 256, device=dev); W = torch.randn(256, 256, device=dev) / 16
nb = C.c_size_t(0); lib.b2a_mlp_packed_bytes(256, 256, C.byref(nb))
Wp = torch.empty(nb.value, dtype=torch.uint8, device=dev)
lib.b2a_mlp_pack_weights(W.data_ptr(), 256, 256, 256, 0, Wp.data_ptr(), Wp.numel(), st)
out = torch.empty(rows, 256, device=dev)
h = C.CDLL(L.LIB_PATH)
mhz = 1965.0          # SM clock of the pool's B200s under load (bench.py clocks.sm_mhz)
for it in range(3):
    lib.b2a_mlp_rows_gemm(A.data_ptr(), 256, rows, 256, Wp.data_ptr(), 256, 1, 3, 0, None, None, None, 0, None, None, out.data_ptr(), 256, st)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 128)(); h.b2a_debug_mlp_trace(buf)
    t = np.array(buf[:], dtype=np.int64); t0 = t[0]
    us = lambda i: (t[i] - t0) / mhz
    print("run %d (us since the CTA started, SM clock %.0f MHz): setup done %.2f" % (it, mhz, us(1)))
    for c in range(8):
        print("  chunk %d: stage free %.2f  staged (+ loads of chunk %d issued) %.2f  barrier %.2f | issuer: weights landed %.2f  MMAs committed %.2f"
              % (c, us(2 + 3 * c), c + 2, us(3 + 3 * c), us(4 + 3 * c), us(64 + 2 * c), us(65 + 2 * c)))
    print("  loop end %.2f  all MMAs done %.2f  epilogue done %.2f  exit barrier %.2f" % (us(40), us(41), us(42), us(43)))
