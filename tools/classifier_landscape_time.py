"""Classifier lml landscape (64 x 64, diffusion) on the C1 grid (30 x 30, 90 training nodes, +-1 labels): default path against the
batched Laplace Newton loop of the spectral path."""
import os, sys, time
import numpy as np
import pandas as pd
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpnet_b200
from oracle import configs as cfg
G, tr, te, t = cfg.grid_case(30, 30, 90, 810)
y = np.where(t > 0, 1, -1).astype(np.int64)
ys = pd.Series(y[tr], index=tr)
th = np.log([0.6, 0.1])
ax1 = np.linspace(np.log(0.01), np.log(0.99), 64); ax2 = np.linspace(np.log(0.001), np.log(10.0), 64)
res = {}
for spec in (True, False):
    m = gpnet_b200.GPnetClassifier(Graph=G, training_nodes=tr, test_nodes=te, theta=list(th), relabel_nodes=True, kerneltype="diffusion",
                                   training_values=ys)
    if spec:
        m.use_spectral()
    n1 = 64 if spec else 8
    m.lml_landscape(list(th), [0, 1], ax1[:n1], ax2)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lml = m.lml_landscape(list(th), [0, 1], ax1[:n1], ax2)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[spec] = lml
    print("spectral=%s: %d points in %.1f ms = %.1f us per point" % (spec, lml.size, dt * 1e3, dt / lml.size * 1e6), file=sys.stderr)
fin = np.isfinite(res[False]) & np.isfinite(res[True][:8])
print("max rel diff on the common rows:", float(np.max(np.abs(res[True][:8][fin] - res[False][fin]) / np.abs(res[False][fin]))), file=sys.stderr)
