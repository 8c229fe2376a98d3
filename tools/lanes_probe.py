"""Throughput of batches of independent logp+grad evaluations at C3 with J concurrent lanes (contexts sharing the chip)."""
import os, sys, time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpnet_b200._lib import Context
from gpnet_b200.engine import evaluate_batch, graph_csr
from oracle import configs as cfg

N = int(os.environ.get("N", 8192))
G, tr, te, t = cfg.rgg_case(N, N // 2, N // 2)
indptr, indices, w, _, _ = graph_csr(G)
B = int(os.environ.get("B", 48))
thetas = [np.log([0.6, 0.01]) + 0.004 * (b - B / 2) / B * np.array([1.0, -1.0]) for b in range(B)]
ref = None
for J in [int(x) for x in os.environ.get("LANES", "1,2,3,4,6").split(",")]:
    for K in [int(x) for x in os.environ.get("PARTS", "0").split(",")]:
        ctxs = [Context(0) for _ in range(J)]
        for c in ctxs:
            c.set_graph(indptr, indices, w); c.set_nodes(tr, te); c.set_targets(t[tr]); c.set_dissection_parts(K)
            bud = os.environ.get('BUDGET')
            c.set_sm_budget(int(bud) if bud is not None else (c.num_sms() // J if J > 1 else 0))
        out, infos = evaluate_batch(ctxs, 1, thetas, None, True)       # warm-up
        torch.cuda.synchronize()
        for c in ctxs: c.reset_counters()
        t0 = time.perf_counter()
        for _ in range(3):
            out, infos = evaluate_batch(ctxs, 1, thetas, None, True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        if ref is None: ref = out
        bad = np.nonzero(np.max(np.abs(out-ref)/np.abs(ref), axis=1) > 1e-11)[0]
        print("  rows off:", bad.tolist(), "cols:", [int(np.argmax(np.abs(out[i]-ref[i])/np.abs(ref[i]))) for i in bad])
        print(f"lanes={J} parts={K} route={ctxs[0].last_route()} evals/s={B/dt:.1f} ms/eval={dt/B*1e3:.3f} max_rel_diff_vs_first={np.max(np.abs(out-ref)/np.abs(ref)):.1e} infos={infos.max()} gflop/eval={sum(c.flop_count() for c in ctxs) / (3 * B) * 1e-9:.2f} layout={ctxs[0].debug_d6_layout() if os.environ.get('LAYOUT') else ''}", flush=True)
        for c in ctxs: c.close()
