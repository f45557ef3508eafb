import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("3danimals_b200._lib"); lib = L.lib()
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
rows, M, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
Pm = torch.randn(rows, M, device=dev); Qm = torch.randn(rows, N, device=dev)
out = torch.zeros(M, N, device=dev)
L.check(lib.b2a_mlp_wgrad(Pm.data_ptr(), M, 0, Qm.data_ptr(), N, 0, rows, M, N, 3, out.data_ptr(), N, 0, st))
torch.cuda.synchronize()
ref = Pm.double().t() @ Qm.double()
print("rel err %.2e" % ((out.double() - ref).abs().max() / ref.abs().max()).item())
