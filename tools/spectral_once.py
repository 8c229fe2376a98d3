"""C1 (30 x 30 grid, n = 90): eigendecomposition + one 64 x 64 landscape through the batched evaluator (ncu target)."""
import os, sys
import numpy as np
import pandas as pd
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpnet_b200
from oracle import configs as cfg
G, tr, te, t = cfg.grid_case(30, 30, 90, 810)
m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=tr, test_nodes=te, theta=list(np.log([0.6, 0.01])), relabel_nodes=True,
                              kerneltype="diffusion", training_values=False)
m.set_training_values(pd.Series(t[tr], index=tr))
m.use_spectral()
ax1 = np.linspace(np.log(0.01), np.log(0.99), 64); ax2 = np.linspace(np.log(0.001), np.log(10.0), 64)
for _ in range(2):
    lml = m.lml_landscape([0.0, 0.0], [0, 1], ax1, ax2)
print("landscape max", float(lml.max()), "jacobi", m._engine.spectral_info)
