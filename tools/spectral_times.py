"""Timings of the spectral path: Jacobi eigendecomposition per N, kernel build vs the default algorithms, C1 / C4 landscapes."""
import os, sys, time, json
import numpy as np
import pandas as pd
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpnet_b200
from gpnet_b200._lib import Context, KERNEL_IDS
from gpnet_b200.engine import graph_csr
from oracle import configs as cfg

out = {}
ctx = Context(0)
for N in [int(x) for x in os.environ.get("NS", "900,2048,4096").split(",")]:
    G = cfg.relabel(__import__("networkx").grid_2d_graph(30, 30)) if N == 900 else cfg.rgg_case(N, N // 2, 0)[0]
    ip, ix, w, _, _ = graph_csr(G)
    ctx.set_graph(ip, ix, w)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.reset_counters()
    sweeps, off = ctx.spectral_prepare()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    fl = ctx.flop_count()
    th = np.log([0.6, 0.01])
    res = {"jacobi_s": dt, "sweeps": sweeps, "off_rel": off, "jacobi_gemm_tflops": fl / dt * 1e-12, "launches": ctx.launch_count()}
    for kt in ("diffusion", "regularized_laplacian"):
        for spec in (False, True):
            ctx.spectral_enable(spec)
            ctx.kernel_build(KERNEL_IDS[kt], th, False)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(3):
                ctx.kernel_build(KERNEL_IDS[kt], th, False)
            ctx.sync(); res[f"kernel_build_ms_{kt}_{'spectral' if spec else 'default'}"] = (time.perf_counter() - t0) / 3 * 1e3
    ctx.spectral_enable(False)
    out[N] = res
    print(N, json.dumps(res), flush=True)
ctx.close()
# C4 landscape (64 x 64, diffusion, N = 4096, n = 2048) default vs spectral, through the public API
G, train, _, t = cfg.rgg_case(4096, 2048, 0)
ax1 = np.linspace(np.log(0.01), np.log(0.99), 64); ax2 = np.linspace(np.log(0.001), np.log(10.0), 64)
for spec in (False, True):
    m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=train, test_nodes=[], theta=list(np.log([0.6, 0.01])), relabel_nodes=True,
                                  kerneltype="diffusion", training_values=False)
    m.set_training_values(pd.Series(t[train], index=train))
    if spec: m.use_spectral()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    l0 = m.lml_landscape([0.0, 0.0], [0, 1], ax1, ax2)
    torch.cuda.synchronize(); first = time.perf_counter() - t0
    t0 = time.perf_counter()
    l1 = m.lml_landscape([0.0, 0.0], [0, 1], ax1, ax2)
    torch.cuda.synchronize(); second = time.perf_counter() - t0
    print("C4 landscape spectral=%s first %.3f s (incl. setup) second %.3f s" % (spec, first, second), flush=True)
    if spec: print("max rel diff vs default:", float(np.max(np.abs(l1 - ref) / np.abs(ref))))
    ref = l1
