"""Small end-to-end pass for compute-sanitizer (memcheck): every kernel family once at tiny sizes."""
import os, sys
import numpy as np, pandas as pd, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpnet_b200
from oracle import configs as cfg
G, tr, te, t = cfg.grid_case(18, 17, 140, 120)          # N = 306: ragged tiles, dataflow potrf with T = 3 (N) and T = 2 (n)
for kt, th in (("diffusion", np.log([0.6, 0.01])), ("regularized_laplacian", np.log([0.6, 0.01])), ("pstep_walk", np.log([2.3, 4, 0.01]))):
    m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=tr, test_nodes=te, theta=list(th), relabel_nodes=True, kerneltype=kt,
                                  training_values=False)
    m.set_training_values(pd.Series(t[tr], index=tr))
    m.predict()
    shapes = (m.k.shape, m.v.shape, m.L.shape)
    print(kt, m.logPosterior_and_grad(th, tr, m.training_values), shapes)
y = np.where(np.sin(0.6 * cfg.bfs_dist(G)) > 0, 1, -1)
c = gpnet_b200.GPnetClassifier(Graph=G, training_nodes=tr, test_nodes=te, theta=list(np.log([0.6, 0.1])), relabel_nodes=True,
                               kerneltype="diffusion", training_values=pd.Series(y[tr], index=tr))
c.predict()
print("classifier", c.newton_iterations, c.gradLogPosterior(c.theta, tr, c.training_values))
print("landscape", m.lml_landscape(list(np.log([2.3, 4, 0.01])), [0, 2], np.log([2.1, 2.6]), np.log([0.01, 0.1])))
torch.cuda.synchronize()
print("done")
