"""Phase timing of route level 6 at C3 (GPNET_B200_D6_TIMING=1 prints CUDA-event times per phase to stderr)."""
import os, sys
os.environ["GPNET_B200_D6_TIMING"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpnet_b200._lib import Context
from oracle import configs as cfg
N = int(os.environ.get("N", 8192))
G, tr, te, t = cfg.rgg_case(N, N // 2, N // 2)
from gpnet_b200.engine import graph_csr
indptr, indices, w, _, _ = graph_csr(G)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = Context(0, stream.cuda_stream)
ctx.set_graph(indptr, indices, w); ctx.set_nodes(tr, te); ctx.set_targets(t[tr])
th = np.log([0.6, 0.01])
for K in [int(x) for x in os.environ.get("PARTS", "4").split(",")]:
    ctx.set_dissection_parts(K)
    for g in (True, False):
        for i in range(3):
            if i == 2: print(f"--- K={K} grad={g} {ctx.last_route()}", file=sys.stderr, flush=True)
            ctx.gpr_logp_grad(1, th, None, g)
