"""Library yardsticks on the GPU box (context only, never the product path):
cuBLAS DGEMM / cuSOLVER DPOTRF via torch, and host-side numpy timings + CPU info."""
import json, os, time, sys
import numpy as np
import torch

def ev_time(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {"cpu_count": os.cpu_count()}
try:
    import cpuinfo
    out["cpu"] = cpuinfo.get_cpu_info().get("brand_raw")
except Exception as e:
    out["cpu"] = str(e)
try:
    from threadpoolctl import threadpool_info
    out["blas"] = [(d.get("internal_api"), d.get("version"), d.get("num_threads")) for d in threadpool_info()]
except Exception as e:
    out["blas"] = str(e)
print(torch.cuda.get_device_name(0))
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    ms = ev_time(lambda: torch.matmul(a, b))
    out[f"cublas_dgemm_{n}_tflops"] = 2 * n**3 / ms * 1e-9
    spd = a @ a.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    ms = ev_time(lambda: torch.linalg.cholesky(spd))
    out[f"cusolver_potrf_{n}_ms"] = ms
    out[f"cusolver_potrf_{n}_tflops"] = n**3 / 3 / ms * 1e-9
    L = torch.linalg.cholesky(spd)
    ms = ev_time(lambda: torch.cholesky_inverse(L))
    out[f"cusolver_potri_{n}_ms"] = ms
    del a, b, spd, L
# host numpy
for n in (2048, 4096):
    a = np.random.rand(n, n)
    t = time.time(); a @ a; out[f"host_dgemm_{n}_gflops"] = 2 * n**3 / (time.time() - t) * 1e-9
print(json.dumps(out, indent=1))
