"""C4: log-marginal-likelihood landscape on a hyper-parameter grid, sharded cyclically over the ranks of a torchrun job
(one process per GPU) with an NCCL all-gather of the (1+P) doubles per point (GPnet.py:539-552).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        tools/landscape_bench.py --nodes 4096 --grid 64 [--check]
"""
import argparse
import contextlib
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=4096)
    ap.add_argument("--grid", type=int, default=64)
    ap.add_argument("--kernel", default="diffusion")
    ap.add_argument("--check", action="store_true", help="compare a few points with the CPU oracle (rank 0)")
    args = ap.parse_args()
    import networkx as nx
    import pandas as pd
    import torch
    import torch.distributed as dist
    import gpnet_b200
    from gpnet_b200.sharded import sharded_landscape

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = args.nodes
    G = nx.random_geometric_graph(N, math.sqrt(20.0 / (math.pi * N)), seed=0)
    perm = np.random.RandomState(0).permutation(N)
    train = sorted(int(x) for x in perm[: N // 2])
    pos = np.array([G.nodes[i]["pos"] for i in range(N)])
    t = np.sin(4 * np.pi * pos[:, 0]) * np.cos(4 * np.pi * pos[:, 1])
    with contextlib.redirect_stdout(sys.stderr):
        m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=train, test_nodes=[x for x in range(N) if x not in set(train)][:8],
                                      theta=list(np.log([0.6, 0.01])), relabel_nodes=True, kerneltype=args.kernel, training_values=False)
    m.set_training_values(pd.Series(t[train], index=train))
    ax1 = np.linspace(np.log(0.01), np.log(0.99), args.grid)
    ax2 = np.linspace(np.log(0.001), np.log(10.0), args.grid)
    sharded_landscape(m, list(m.theta), [0, 1], ax1[:2], ax2[:2])          # warm-up (allocations, NCCL init)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lml = sharded_landscape(m, list(m.theta), [0, 1], ax1, ax2)[..., 0]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dt = float(tmax.item())
        chk = torch.from_numpy(lml.copy()).cuda()
        ref = chk.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(chk, ref), "ranks disagree on the gathered landscape"
    if rank == 0:
        out = {"workload": f"C4 landscape {args.grid}x{args.grid}, {args.kernel}, RGG N={N}, n={N//2}", "n_gpus": world,
               "seconds": dt, "points_per_s": args.grid ** 2 / dt, "lml_max": float(np.max(lml)),
               "argmax": [int(x) for x in np.unravel_index(np.argmax(lml), lml.shape)], "finite": bool(np.all(np.isfinite(lml)))}
        if args.check:
            from oracle import gp_oracle as go
            orc = go.OracleGP(G, kerneltype=args.kernel, expm_mode="dense")
            errs = []
            for (i, j) in ((0, 0), (args.grid // 2, args.grid // 3), (args.grid - 1, args.grid - 1)):
                ref = -orc.neg_logp([ax1[i], ax2[j]], train, t[train])
                errs.append(abs(ref - lml[i, j]) / abs(ref))
            out["oracle_relerr_3pts"] = errs
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
