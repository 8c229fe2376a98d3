"""Does the TMA GEMM give bit-identical results when other kernels share the GPU?  (debugging aid)
NOISE=transpose: a strided torch copy on another stream."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpnet_b200._lib import Context
ctx = Context(0)
torch.manual_seed(0)
M = N = int(os.environ.get("MN", 870)); K = int(os.environ.get("K", 800))
ld = 8192
A = torch.randn(M + 64, ld, dtype=torch.float64, device="cuda")
ldc = (N + 127) // 128 * 128
C = torch.zeros(ldc, ldc, dtype=torch.float64, device="cuda")
noise_stream = torch.cuda.Stream()
big = torch.randn(2048, 65536, dtype=torch.float64, device="cuda")       # 1 GB
MODE = os.environ.get("NOISE", "transpose")
torch.cuda.synchronize()
def run(noise):
    C.zero_()
    torch.cuda.synchronize()
    if noise:
        with torch.cuda.stream(noise_stream):
            for _ in range(4):
                y = big.t().contiguous() if MODE == "transpose" else torch.sin(big)
    ctx.gemm(0, 1, M, N, K, 1.0, A.data_ptr() + 130 * 8, ld, A.data_ptr() + 130 * 8, ld, 0.0, C.data_ptr(), ldc)
    ctx.sync()
    torch.cuda.synchronize()
    return C.clone()
ref = run(False)
ok = (A[:M, 130:130 + K] @ A[:M, 130:130 + K].T)
print("vs torch:", float((ref[:M, :N] - ok).abs().max()))
nbad = 0
for it in range(int(os.environ.get('ITERS', 12))):
    out = run(True)
    d = (out - ref).abs().cpu().numpy()
    n = int((d > 0).sum())
    if n:
        nbad += 1
        bad = np.argwhere(d > 0)
        tiles = sorted(set((int(r) // 64, int(c) // 64) for r, c in bad))
        r0, c0 = bad[0]
        t0 = (r0 // 64, c0 // 64)
        inb = bad[(bad[:, 0] // 64 == t0[0]) & (bad[:, 1] // 64 == t0[1])]
        print(it, "wrong entries", n, "max", d.max(), "tiles(64):", tiles[:10], "...", len(tiles),
              "| in first tile: rows", sorted(set(inb[:, 0] % 64))[:16], "cols", sorted(set(inb[:, 1] % 64))[:16],
              "| value got/ref", float(out[r0, c0]), float(ref[r0, c0]), flush=True)
print("iterations with wrong results:", nbad)
