"""Development probe for route level 6 (whole-graph dissection): parity against the other routes and timing at C3 for
several part counts.  Run on the GPU box: python tools/d6_probe.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpnet_b200._lib import Context
from oracle import configs as cfg

N = int(os.environ.get("N", 8192))
indptr, indices, w, pos, t = cfg.rgg_csr_fast(N)
tr, te = cfg.split(N, N // 2, N // 2)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = Context(0, stream.cuda_stream)
ctx.set_graph(indptr, indices, w); ctx.set_nodes(tr, te); ctx.set_targets(t[tr])
th = np.log([0.6, 0.01])
ctx.set_fast_paths(5)
ref = ctx.gpr_logp_grad(1, th, None, True)[1]
refv = ctx.gpr_logp_grad(1, th, None, False)[1]
print("level5", ref, ctx.last_route())
for K in [int(x) for x in os.environ.get("PARTS", "0,2,3,4,5,6,8").split(",")]:
    ctx.set_dissection_parts(K); ctx.set_fast_paths(6)
    for g in (True, False):
        st, out, info = ctx.gpr_logp_grad(1, th, None, g)
        route = ctx.last_route()
        for _ in range(3): ctx.gpr_logp_grad(1, th, None, g)
        ctx.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20): ctx.gpr_logp_grad(1, th, None, g)
        e1.record(stream); torch.cuda.synchronize()
        r = ref if g else refv
        print(f"K={K} grad={g} route={route} st={st} info={info} ms={e0.elapsed_time(e1)/20:.3f} launches={ctx.launch_count()/20:.0f} "
              f"gflop={ctx.flop_count()/20*1e-9:.1f} relerr={np.max(np.abs(out[:len(r)]-r)/np.abs(r)):.2e}", flush=True)
