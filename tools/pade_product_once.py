"""One Pade product of the N = 8192 expm as the GEMM kernel sees it: C = A A (symmetric A = -beta L~ of the C3 graph; lower tiles +
mirrored store, 128x128x16 tiles).  gpn_syrk_f64 issues exactly that launch (A A^T with the same flags).  ncu target."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpnet_b200._lib import Context
from bench import c3_problem
N = 8192
G, csr, train, test, t = c3_problem(N)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = Context(0, stream.cuda_stream)
ctx.set_graph(*csr)
A = torch.zeros(N, N, dtype=torch.float64, device="cuda"); C = torch.zeros_like(A)
ctx.laplacian_dense(-0.6, 0.0, A.data_ptr(), N)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    e0.record(stream)
    ctx.syrk(N, N, 1.0, A.data_ptr(), N, 0.0, C.data_ptr(), N)
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
fl = 2.0 * 128 * 128 * N * (64 * 65 // 2)      # 2080 lower tiles of 128 x 128, K = N
print("A^2 at N=8192 (lower tiles + mirror): %.3f ms, executed %.1f GFLOP, %.2f TFLOP/s" % (ms, fl * 1e-9, fl / (ms * 1e-3) * 1e-12))
ref = (A[:256] @ A)[:, :256]
print("check vs torch:", float((C[:256, :256] - ref).abs().max()))
