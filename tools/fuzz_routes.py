"""One-off randomized sweep: every route of the regularized-Laplacian regressor likelihood against the numpy oracle on random
geometric graphs of random size, random training fraction, random (beta, noise) -- ragged block sizes on both sides of the
256-row threshold of the augmented factorisations."""
import os
import sys

import networkx as nx
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpnet_b200._lib import Context  # noqa: E402
from oracle import gp_oracle as go  # noqa: E402

ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
ctx = Context(0, ts.cuda_stream)
rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
worst = 0.0
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
    N = int(rng.randint(300, 4200))
    frac = rng.uniform(0.12, 0.88)
    n = max(2, int(N * frac))
    beta = float(np.exp(rng.uniform(np.log(0.01), np.log(5.0))))
    noise = float(np.exp(rng.uniform(np.log(1e-4), np.log(10.0))))
    G = nx.random_geometric_graph(N, np.sqrt(rng.uniform(6, 24) / (np.pi * N)), seed=int(rng.randint(1 << 30)))
    perm = rng.permutation(N)
    tr = sorted(int(x) for x in perm[:n])
    te = sorted(int(x) for x in perm[n:n + min(N - n, 16)])
    t = rng.randn(n)
    indptr, indices, w, _ = go.graph_to_csr(G)
    ctx.set_graph(indptr, indices, w)
    ctx.set_nodes(tr, te)
    th = np.log([beta, noise])
    ref_v, ref_g = go.OracleGP(G, kerneltype="regularized_laplacian").neg_logp_grad(th, tr, t)
    errs = []
    for level in (0, 1, 2, 3, 4, 5):
        ctx.set_fast_paths(level)
        st, out, info = ctx.gpr_logp_grad(1, th, t, True)
        assert st == 0 and info == 0, (case, level, st, info)
        ev = abs(out[0] - ref_v) / abs(ref_v)
        eg = float(np.max(np.abs(out[1:] - ref_g)) / np.max(np.abs(ref_g)))
        errs.append(max(ev, eg))
    worst = max(worst, max(errs))
    print(f"case {case:2d}: N={N:4d} n={n:4d} o={N-n:4d} beta={beta:.3g} noise={noise:.3g}  max rel err per route {['%.1e' % e for e in errs]}", flush=True)
print("worst", worst)
assert worst < 1e-8
