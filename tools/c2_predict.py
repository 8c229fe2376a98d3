"""C2 (Barabasi-Albert N = 2000, n = m = 1000, classifier): time of predict(); run under ncu for the launch list."""
import os, sys, time
import numpy as np
import pandas as pd
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpnet_b200
from oracle import configs as cfg
kt = os.environ.get("KT", "diffusion")
th = np.log([0.6, 0.1]) if kt != "pstep_walk" else np.log([2.3, 4, 0.1])
G, tr, te, y = cfg.ba_case(2000, 1000, 1000)
m = gpnet_b200.GPnetClassifier(Graph=G, training_nodes=tr, test_nodes=te, theta=list(th), relabel_nodes=True, kerneltype=kt,
                               training_values=pd.Series(y[tr], index=tr))
if os.environ.get("SPECTRAL"):
    m.use_spectral()
reps = int(os.environ.get("REPS", 5))
m.predict()
torch.cuda.synchronize()
ctx = m._engine.ctx
ctx.reset_counters()
t0 = time.perf_counter()
for _ in range(reps):
    m._engine.invalidate_kernel()
    m.predict()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print(f"C2 {kt} spectral={bool(os.environ.get('SPECTRAL'))}: predict {dt*1e3:.2f} ms, Newton iterations {m.newton_iterations}, launches per predict {ctx.launch_count()/reps:.0f}", file=sys.stderr)
