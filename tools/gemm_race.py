"""Does the TMA GEMM give bit-identical results when other kernels share the GPU?  (debugging aid)
NOISE=transpose: a strided torch copy on another stream.  GPNET_B200_GEMM_DEBUG selects kernel variants (common.cuh GF_DBG_*)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpnet_b200 import _lib
from gpnet_b200._lib import Context
if os.environ.get('GEMM_RACE_LIB'):
    _lib.load_library(os.path.abspath(os.environ['GEMM_RACE_LIB']))      # e.g. a build with -DGEMM_SMALL_MIN_CTAS=1 (the reproducer)
ctx = Context(0)
K = int(os.environ.get("K", 800))
OFF = int(os.environ.get("OFF", 130))
ld = 8192
MODE = os.environ.get("NOISE", "transpose")
ITERS = int(os.environ.get("ITERS", 12))
noise_stream = torch.cuda.Stream()
big = torch.randn(2048, 32768, dtype=torch.float64, device="cuda")       # 512 MB
for MN in [int(x) for x in os.environ.get("MN", "300,870,896").split(",")]:
    M = N = MN
    torch.manual_seed(MN)
    A = torch.randn(M + 64, ld, dtype=torch.float64, device="cuda")
    ldc = (N + 127) // 128 * 128
    C = torch.zeros(ldc, ldc, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def run(noise):
        C.zero_()
        torch.cuda.synchronize()
        if noise:
            with torch.cuda.stream(noise_stream):
                for _ in range(4):
                    y = big.t().contiguous() if MODE == "transpose" else torch.sin(big)
        ctx.gemm(0, 1, M, N, K, 1.0, A.data_ptr() + OFF * 8, ld, A.data_ptr() + OFF * 8, ld, 0.0, C.data_ptr(), ldc)
        ctx.sync()
        torch.cuda.synchronize()
        return C.clone()
    ref = run(False)
    ok = (A[:M, OFF:OFF + K] @ A[:M, OFF:OFF + K].T)
    err0 = float((ref[:M, :N] - ok).abs().max())
    nbad = 0
    for it in range(ITERS):
        out = run(True)
        d = (out - ref).abs().cpu().numpy()
        n = int((d > 0).sum())
        if n:
            nbad += 1
            if nbad <= 3:
                bad = np.argwhere(d > 0)
                tiles = sorted(set((int(r) // 64, int(c) // 64) for r, c in bad))
                r0, c0 = bad[0]
                t0 = (r0 // 64, c0 // 64)
                inb = bad[(bad[:, 0] // 64 == t0[0]) & (bad[:, 1] // 64 == t0[1])]
                print("  ", it, "wrong entries", n, "max", d.max(), "tiles(64):", tiles[:10], "...", len(tiles),
                      "| first tile: rows", sorted(set(int(x) for x in inb[:, 0] % 64)), "cols", sorted(set(int(x) for x in inb[:, 1] % 64)),
                      "| got/ref", float(out[r0, c0]), float(ref[r0, c0]), flush=True)
    print(f"DEBUG={os.environ.get('GPNET_B200_GEMM_DEBUG', '0')} OFF={OFF} MN={MN} vs_torch={err0:.1e} bad_iterations={nbad}/{ITERS}", flush=True)
