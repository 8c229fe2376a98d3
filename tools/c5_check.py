"""C5: large-N single-GPU diffusion-kernel GPR on the 128x256 grid graph (N = 32768, n = m = 16384; ~8.6 GB per N x N
fp64 matrix).  Times one logp, one logp+grad and one predict, and checks the semigroup identity
exp(-b L~) exp(-b L~) = exp(-2b L~) on a column sample (device-side product through the C-ABI GEMM)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpnet_b200._lib import Context  # noqa: E402
from oracle import configs as cfg  # noqa: E402

a, b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 256)
N = a * b
n = N // 2
indptr, indices, w = cfg.grid_csr_fast(a, b)
dist = (np.arange(N) // b + np.arange(N) % b).astype(np.float64)       # BFS distance from node 0 on a grid
t = np.sin(0.05 * dist)
train, test = cfg.split(N, n, n)
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
ctx = Context(0, ts.cuda_stream)
ctx.set_graph(indptr, indices, w)
ctx.set_nodes(train, test)
out = {"N": N, "n": n}
th = np.log([0.6, 0.01])


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, time.perf_counter() - t0


(_, o, info), dt = timed(lambda: ctx.gpr_logp_grad(0, th, t[train], False))
out["logp_first_call_s"] = dt
(_, o, info), dt = timed(lambda: ctx.gpr_logp_grad(0, th, t[train], False))
out.update(logp=-o[0], logp_s=dt, info=info, pade=ctx.expm_info())
(_, og, info), dt = timed(lambda: ctx.gpr_logp_grad(0, th, t[train], True))
out.update(logp_grad_s=dt, grad=list(og[1:]), logp_consistent=abs(og[0] - o[0]) / abs(o[0]))
(_, pr), dt = timed(lambda: ctx.gpr_predict(0, th, t[train]))
out.update(predict_s=dt, fstar_rmse=float(np.sqrt(np.mean((pr["fstar"] - t[test]) ** 2))), s_mean=float(pr["s"].mean()),
           infos=[pr["info_k"], pr["info_kss"]])
# semigroup identity on 64 columns
cols = np.arange(0, N, N // 64, dtype=np.int32)[:64]
ctx.kernel_build(0, np.log([0.3, 0.01]), False)
ptr, ld = ctx.kernel_ptr()
K1 = torch.empty((N, ld), dtype=torch.float64, device="cuda")
import ctypes
torch.cuda.synchronize()
ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(K1.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(N * ld * 8), 3)
K1c = torch.zeros((N, 64), dtype=torch.float64, device="cuda")
K1c.copy_(K1[:, torch.from_numpy(cols.astype(np.int64)).cuda()])
P = torch.zeros((N, 64), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
ctx.gemm(0, 0, N, 64, N, 1.0, K1.data_ptr(), ld, K1c.data_ptr(), 64, 0.0, P.data_ptr(), 64)
ctx.sync()
ctx.kernel_build(0, np.log([0.6, 0.01]), False)
K2c = ctx.kernel_block(range(N), cols, 0.0)
out["semigroup_maxabs"] = float(np.max(np.abs(P.cpu().numpy() - K2c)))
out["mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
free, total = torch.cuda.mem_get_info()
out["device_mem_used_GB"] = (total - free) / 1e9
print(json.dumps(out))
