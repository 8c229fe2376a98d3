"""Wall-clock of the non-headline configurations through the drop-in classes on one GPU, next to the
CPU oracle (dense restatement) on the host: C1 regressor (30x30 grid), C2 classifier (BA 2000), one C4 landscape point
(RGG 4096, diffusion), C3 predict."""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpnet_b200  # noqa: E402
from oracle import configs as cfg  # noqa: E402
from oracle import gp_oracle as go  # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def best(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


out = {}
L = np.log
# ---- C1
G, tr, te, t = cfg.grid_case(30, 30, 90, 810)
for kt, th in (("diffusion", L([0.6, 0.01])), ("regularized_laplacian", L([0.6, 0.01])), ("pstep_walk", L([2.3, 4, 0.01]))):
    m = quiet(gpnet_b200.GPnetRegressor, Graph=G, training_nodes=tr, test_nodes=te, theta=list(th), relabel_nodes=True, kerneltype=kt,
              training_values=False)
    m.set_training_values(pd.Series(t[tr], index=tr))
    orc = go.OracleGP(G, kerneltype=kt, expm_mode="dense")
    out[f"C1_{kt}"] = {"gpu_predict_ms": 1e3 * best(lambda: quiet(m.predict)), "gpu_logp_ms": 1e3 * best(lambda: m.logp()),
                       "gpu_logp_grad_ms": 1e3 * best(lambda: m.logPosterior_and_grad(th, tr, m.training_values)),
                       "cpu_oracle_predict_ms": 1e3 * best(lambda: orc.predict_regressor(th, tr, te, t[tr]), 2),
                       "cpu_oracle_logp_ms": 1e3 * best(lambda: orc.neg_logp(th, tr, t[tr]), 2)}
# ---- C2
G, tr, te, y = cfg.ba_case(2000, 1000, 1000)
for kt, th in (("diffusion", L([0.6, 0.1])), ("pstep_walk", L([2.1, 2, 0.1]))):
    c = quiet(gpnet_b200.GPnetClassifier, Graph=G, training_nodes=tr, test_nodes=te, theta=list(th), relabel_nodes=True, kerneltype=kt,
              training_values=pd.Series(y[tr], index=tr))
    orc = go.OracleGP(G, kerneltype=kt, expm_mode="dense")
    out[f"C2_{kt}"] = {"gpu_predict_ms": 1e3 * best(lambda: quiet(c.predict)), "newton_iterations": c.newton_iterations,
                       "cpu_oracle_predict_ms": 1e3 * best(lambda: orc.predict_classifier(th, tr, te, y[tr]), 1)}
# ---- C4 single landscape point (diffusion, N = 4096, n = 2048) at three betas (Pade orders differ)
G, tr, te, t = cfg.rgg_case(4096, 2048, 8)
m = quiet(gpnet_b200.GPnetRegressor, Graph=G, training_nodes=tr, test_nodes=te, theta=list(L([0.6, 0.01])), relabel_nodes=True,
          kerneltype="diffusion", training_values=False)
m.set_training_values(pd.Series(t[tr], index=tr))
orc = go.OracleGP(G, kerneltype="diffusion", expm_mode="dense")
for beta in (0.02, 0.3, 0.9):
    th = L([beta, 0.01])
    out[f"C4_point_beta{beta}"] = {"gpu_logp_ms": 1e3 * best(lambda: m.logPosterior(th, tr, m.training_values)),
                                   "pade": m._engine.ctx.expm_info(),
                                   "cpu_oracle_logp_ms": 1e3 * best(lambda: orc.neg_logp(th, tr, t[tr]), 1)}
print(json.dumps(out, indent=1))
