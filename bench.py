"""Headline benchmark: GP logp+grad evaluations per second at N = 8192 graph nodes (fp64), config C3 of SURVEY.md
Appendix D (GPnetRegressor, regularized-Laplacian kernel, random geometric graph, n = 4096 training nodes, P = 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c4] [--evals-per-step B]

One "step" = one batch of B multi-start evaluations theta (host) -> (-logp, -dlogp/dtheta) (host), each from scratch: nothing
is carried over between evaluations (like the reference, which rebuilds everything per call); only the graph, the node sets
and the targets are resident.  `value` = evaluations per second over all ranks (weak scaling: every GPU runs its own batch;
the (1+P)-double results are all-gathered once, after the K steps, inside the timed region).  The default route never forms
the kernel; `routes` times the reference's own order of operations (full kernel, block gather, Cholesky,
alpha, log-det, K_y^-1, the P trace terms) beside it, and utilisation is always reported from EXECUTED flops.

`--workload c4` makes the 64 x 64 diffusion landscape at N = 4096 (config C4, GPnet.py:539-552; strong scaling over the ranks,
kernel groups dealt longest-first, ONE all-gather) the headline; in the default run it is reported under `extras.c4_landscape`
at every N so that the driver's scaling runs carry it.  Prints ONE JSON line; rank 0 only.
"""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # before CUDA is initialised: see gpnet_b200/__init__.py

if "--impl" in sys.argv and "reference" in sys.argv:
    # the CPU arm runs single-process on ALL host cores: torchrun exports OMP_NUM_THREADS=1, undo that before numpy loads its BLAS
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.pop(_k, None)

import argparse
import contextlib
import json
import math
import subprocess
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GP logp+grad evals/sec at N=8192 nodes (fp64)"
UNIT = "evals/s"
THETA = np.log([0.6, 0.01])      # C3 theta_0
KERNEL = "regularized_laplacian"

_JSON_FD = None


def emit(line):
    """The one JSON line of the contract, on the original stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def workload_config(workload, N, n, P, world, B):
    """The same `config` object on both arms."""
    if workload == "c4":
        return {"workload": f"C4: GPnetRegressor diffusion, 64x64 logp landscape (theta0 x noise) on a random geometric graph N={N}, n={n}; "
                            "one whole landscape per step, grid sharded over the GPUs by kernel group, ONE all-gather of (1+P) doubles per point",
                "l2": "inputs larger than L2 (each N x N fp64 operand is 134 MB vs 126 MB L2; ten are live per kernel build)",
                "parallelism": f"grid sharded x{world}"}
    return {"workload": f"C3: GPnetRegressor regularized_laplacian, random geometric graph N={N}, n={n}, P={P}, theta=log[0.6,0.01]; "
                        f"one step = {B} multi-start logp+grad evaluations per GPU, each from scratch; (1+P)-double results all-gathered "
                        "once after the K steps (inside the timed region)",
            "evals_per_step": B,
            "l2": "inputs larger than L2 (each N x N fp64 operand is 537 MB vs 126 MB L2)", "parallelism": f"replicas x{world}"}


def c3_problem(N):
    """C3: nx.random_geometric_graph(N, r = sqrt(20/(pi N)), seed=0), 50/50 split, t = sin(4 pi x) cos(4 pi y)."""
    from gpnet_b200.engine import graph_csr
    import networkx as nx
    r = math.sqrt(20.0 / (math.pi * N))
    G = nx.random_geometric_graph(N, r, seed=0)
    indptr, indices, w, nodelist, _ = graph_csr(G)
    perm = np.random.RandomState(0).permutation(N)
    n = N // 2
    train = sorted(int(x) for x in perm[:n])
    test = sorted(int(x) for x in perm[n:2 * n])
    pos = np.array([G.nodes[i]["pos"] for i in range(N)])
    t = np.sin(4 * np.pi * pos[:, 0]) * np.cos(4 * np.pi * pos[:, 1])
    return G, (indptr, indices, w), train, test, t


def multistart_thetas(rank, B):
    """B log-parameter vectors around the C3 theta_0 (multi-start / line-search probes of an optimiser), distinct per rank."""
    d = np.array([1.0, -1.0])
    return [THETA + (0.01 * rank + 0.004 * (b - B / 2) / max(B, 1)) * d for b in range(B)]


def algorithmic_flops(N, n):
    """kernel N^3 (POTRF + POTRI) + gradient SYRK n^2 N + block POTRF + POTRI n^3."""
    return float(N) ** 3 + float(n) ** 2 * N + float(n) ** 3


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): one streaming
    `nvidia-smi -lms 50` process, started right before the region and stopped right after it."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)      # let the first sample arrive before the region starts
        except Exception:
            self.proc = None

    def summary(self):
        samples = []
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            samples = [[x.strip() for x in line.split(",")] for line in out.strip().splitlines() if line.strip()]

        def num(s):
            try:
                return float(s)
            except ValueError:
                return None
        sm = [num(s[0]) for s in samples if len(s) >= 7 and num(s[0]) is not None]
        mx = [num(s[1]) for s in samples if len(s) >= 7 and num(s[1]) is not None]
        pw = [num(s[2]) for s in samples if len(s) >= 7 and num(s[2]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in samples if len(s) >= 7 for i in range(4) if s[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the ONLY place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def pin_blas_threads():
    """All host cores for the BLAS/LAPACK pools, also under torchrun (which exports OMP_NUM_THREADS=1).  Returns the count."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=cores)
        return max([d.get("num_threads", 1) for d in threadpool_info()] + [1])
    except Exception:
        return cores


def cpu_port_eval(problem, N):
    """Best-effort dense numpy/scipy restatement with the analytic gradient: ONE full logp+grad (cpu_baseline, kind "port")."""
    from oracle import gp_oracle as go
    _, csr, train, test, t = problem
    orc = go.OracleGP(csr=csr, kerneltype=KERNEL, expm_mode="dense")
    t0 = time.perf_counter()
    val, _ = orc.neg_logp_grad(THETA, train, t[train])
    return time.perf_counter() - t0, val


class ReferenceModel:
    """The reference's own GPnetRegressor (unmodified, under the shims of oracle/ref_shim.py) when its modules are available
    (oracle/_ref on the GPU box, /root/reference in the build container); else the numpy port with the reference's call
    pattern.  logp+grad costs (P+1) logPosterior calls there: optimize_params passes no jac, scipy takes forward differences
    (GPnet.py:264-271)."""

    def __init__(self, problem, N):
        import pandas as pd
        G, csr, train, test, t = problem
        self.train, self.tv = train, t[train]
        self.kind = "port"
        self.model = None
        try:
            from oracle import ref_shim
            path = ref_shim.find_reference()
            if path is not None:
                import io
                _, REG, _ = ref_shim.load_reference(path)
                with contextlib.redirect_stdout(io.StringIO()):
                    self.model = REG(Graph=G, training_nodes=list(train), test_nodes=list(test), theta=list(THETA), relabel_nodes=True,
                                     kerneltype=KERNEL, training_values=False)
                self.series = pd.Series(self.tv, index=train)
                self.model.set_training_values(self.series.copy())
                self.kind = "reference"
        except Exception as e:      # stay runnable: the port restates the same call pattern
            print("reference modules not usable (%s: %s); timing the numpy port" % (type(e).__name__, e), file=sys.stderr)
            self.model = None
        if self.model is None:
            from oracle import gp_oracle as go
            self.orc = go.OracleGP(csr=csr, kerneltype=KERNEL, expm_mode="dense")

    def log_posterior(self, theta):
        if self.model is not None:
            return float(self.model.logPosterior(np.array(theta), self.train, self.series))      # GPnetRegressor.py:136-150
        return float(self.orc.neg_logp(theta, self.train, self.tv))

    def logp_and_fd_grad(self, theta):
        """What one gradient evaluation costs the reference's optimiser: f(x) and P forward-difference probes."""
        h = 1.4901161193847656e-08      # scipy's default 2-point step sqrt(eps)
        f0 = self.log_posterior(theta)
        g = np.zeros(len(theta))
        for j in range(len(theta)):
            tp = np.array(theta, dtype=float)
            tp[j] += h
            g[j] = (self.log_posterior(tp) - f0) / h
        return f0, g


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, single process (rank 0 alone
    under torchrun; the CPU arm does not scale with the number of GPUs, so `value` is the same at every N).  Every step runs ONE
    full-size logp+grad = all (P+1) logPosterior probes; nothing is extrapolated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = pin_blas_threads()
    N = args.nodes
    problem = c3_problem(N)
    P = len(THETA)
    ref = ReferenceModel(problem, N)
    W, K = args.warmup, args.steps
    thetas = multistart_thetas(0, max(W + K, 1))
    times, val = [], None
    budget = 1500.0                  # safety net: report the steps actually run if the box is much slower than expected
    t_begin = time.perf_counter()
    for i in range(W + K):
        t0 = time.perf_counter()
        val, _ = ref.logp_and_fd_grad(thetas[i % len(thetas)])
        if i >= W:
            times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget and times:
            break
    t_step = float(np.mean(times))
    value = 1.0 / t_step
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": W, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config("c3", N, N // 2, P, args.gpus, args.evals_per_step),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": ref.kind,
                             "sample": f"{len(times)} steps of ONE full-size logp+grad (N={N}, n={N // 2}) = (P+1)={P + 1} logPosterior "
                                       "calls each (forward differences, as the reference's optimize_params gets its gradient); the "
                                       "unmodified GPnetRegressor under oracle/ref_shim.py" if ref.kind == "reference" else
                                       f"{len(times)} steps of (P+1) logPosterior calls of the numpy port (reference modules not found)",
                             "single_process": True},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def recorded_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture committed under profiles/
    (profiles/r02_traffic.json is written by tools/ncu_summary.py from that capture); None when no capture is recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as fh:
            rec = json.load(fh)
        for k, v in rec.items():
            if k != "source" and k in kernel_substr:
                return v, rec.get("source")
    except Exception:
        pass
    return None, None


def c4_problem(N=4096):
    from oracle import configs as cfg      # canonical synthetic inputs only (pure host code, no numerics)
    G, train, _, t = cfg.rgg_case(N, N // 2, 0)
    ax1 = np.linspace(np.log(0.01), np.log(0.99), 64)
    ax2 = np.linspace(np.log(0.001), np.log(10.0), 64)
    return G, train, t, ax1, ax2


def run_c4(ctx, stream, world, rank, local, reps, barrier, spectral=False, first_out=None):
    """The 64 x 64 diffusion landscape of config C4 through the public API (GPnetRegressor.lml_landscape, GPnet.py:539-552):
    host buffers in, host array out, grid sharded over the ranks, one all-gather.  Returns (seconds per landscape max over
    ranks, lml array, device-event ms per landscape)."""
    import pandas as pd
    import torch
    import torch.distributed as dist
    import gpnet_b200
    from gpnet_b200.engine import register_context
    G, train, t, ax1, ax2 = c4_problem()
    with contextlib.redirect_stdout(sys.stderr):
        model = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=train, test_nodes=[], theta=list(np.log([0.6, 0.01])), relabel_nodes=True,
                                          kerneltype="diffusion", training_values=False)
    model.set_training_values(pd.Series(t[train], index=train))
    if spectral:
        model.use_spectral()       # kernels as V r(lam) V^T; the first landscape pays for the eigendecomposition of L~
    register_context(ctx)
    model._engine.device = local
    best, best_ev, lml = None, None, None
    with contextlib.redirect_stdout(sys.stderr):
        for rep in range(reps + 1):            # first repetition = warm-up (allocations, graph upload)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            lml = model.lml_landscape([0.0, 0.0], [0, 1], ax1, ax2)
            e1.record(stream)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            ev = e0.elapsed_time(e1)
            if world > 1:
                tm = torch.tensor([dt, ev], dtype=torch.float64, device="cuda")
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                dt, ev = float(tm[0].item()), float(tm[1].item())
            if rep == 0 and first_out is not None:
                first_out.append(dt)
            if rep > 0 and (best is None or dt < best):
                best, best_ev = dt, ev
    if spectral:
        first_out.append(model._engine.spectral_info)
        model.use_spectral(False)
        model._engine.bind()
    return best, lml, best_ev


def c1_batched_landscapes(ctx, local):
    import pandas as pd
    import torch
    import gpnet_b200
    from gpnet_b200.engine import register_context
    from oracle import configs as cfg      # canonical synthetic inputs only
    G, tr, te, t = cfg.grid_case(30, 30, 90, 810)
    ax1 = np.linspace(np.log(0.01), np.log(0.99), 64)
    ax2 = np.linspace(np.log(0.001), np.log(10.0), 64)
    out = {"what": "C1: 30x30 grid, n = 90, diffusion kernel, 64x64 landscape over (theta0, noise) through lml_landscape with "
                   "use_spectral(): regressor = gpn_gpr_logp_grad_batch, classifier = gpn_gpc_newton_batch (whole Laplace Newton loop per "
                   "CTA); wall clock of the second landscape (the first pays for the eigendecomposition of L~)"}
    with contextlib.redirect_stdout(sys.stderr):
        reg = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=tr, test_nodes=te, theta=list(np.log([0.6, 0.01])), relabel_nodes=True,
                                        kerneltype="diffusion", training_values=False)
        reg.set_training_values(pd.Series(t[tr], index=tr))
        y = np.where(t > 0, 1, -1).astype(np.int64)
        cls = gpnet_b200.GPnetClassifier(Graph=G, training_nodes=tr, test_nodes=te, theta=list(np.log([0.6, 0.1])), relabel_nodes=True,
                                         kerneltype="diffusion", training_values=pd.Series(y[tr], index=tr))
        register_context(ctx)
        for name, m in (("regressor", reg), ("classifier", cls)):
            m._engine.device = local
            m.use_spectral()
            m.lml_landscape([0.0, 0.0], [0, 1], ax1, ax2)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lml = m.lml_landscape([0.0, 0.0], [0, 1], ax1, ax2)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out[name] = {"landscape_ms": dt * 1e3, "us_per_point": dt / lml.size * 1e6, "points": int(lml.size),
                         "finite_points": int(np.isfinite(lml).sum()), "jacobi_sweeps_and_offdiag_rel": list(m._engine.spectral_info or ())}
            m.use_spectral(False)
            m._engine.bind()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--workload", default="c3", choices=["c3", "c4"])
    ap.add_argument("--evals-per-step", type=int, default=120)
    ap.add_argument("--lanes", type=int, default=3, help="evaluations in flight per GPU (contexts sharing the SMs)")
    ap.add_argument("--nodes", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 themselves (NCCL's version banner, the
    # drop-in classes' reference-style prints) are sent to stderr, and the line is written to the saved descriptor at the end
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from gpnet_b200._lib import KERNEL_IDS, Context
    import gpnet_b200
    import pandas as pd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3)
    K = args.steps
    B = max(args.evals_per_step, 1)

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = Context(local, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "c4":
        return main_c4(args, ctx, stream, world, rank, local, W, K, barrier)

    N = args.nodes
    n = N // 2
    problem = c3_problem(N)
    G, csr, train, test, t = problem
    P = len(THETA)
    kind = KERNEL_IDS[KERNEL]
    from gpnet_b200.engine import evaluate_batch, lane_sm_budget
    J = max(args.lanes, 1)
    # J lanes = J contexts bound to the same graph, each with an SM budget for its co-resident factorisation kernels: one
    # evaluation is bound by its dependency chains and leaves SMs idle, three in flight fill them (measured: 1 / 2 / 3 / 4 lanes:
    # 569 / 733 / 820 / 784 evals/s; engine.lane_sm_budget has the budget sweep)
    ctxs = [ctx] + [Context(local) for _ in range(J - 1)]
    for cx in ctxs:
        cx.set_graph(*csr)
        cx.set_nodes(train, test)
        cx.set_targets(t[train])
        cx.set_sm_budget(lane_sm_budget(ctx.num_sms(), J))
    thetas = multistart_thetas(rank, B)
    results = np.zeros((K, B, 1 + P))
    gathered = torch.zeros((world, K * B * (1 + P)), dtype=torch.float64, device="cuda") if world > 1 else None
    health = {"cholesky_info_max": 0}

    def step(k):
        out_b, infos = evaluate_batch(ctxs, kind, thetas, None, True)
        health["cholesky_info_max"] = max(health["cholesky_info_max"], int(infos.max()))
        results[k % K] = out_b

    # the first evaluation of a node set pays for the host-side analysis of the route (BFS level sets, strips x cells dissection,
    # permuted CSR, separator index; once per graph and node set) and for the buffer allocations: timed separately
    t_first = time.perf_counter()
    st0, _, _ = ctx.gpr_logp_grad(kind, thetas[0], None, True)
    first_eval_ms = (time.perf_counter() - t_first) * 1e3
    t_second = time.perf_counter()
    ctx.gpr_logp_grad(kind, thetas[0], None, True)
    second_eval_ms = (time.perf_counter() - t_second) * 1e3
    for k in range(W):
        step(k)
    route, parts = ctx.last_route()
    barrier()
    peak_tf = ctx.fp64_peak(0.3) if rank == 0 else 0.0
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    for cx in ctxs:
        cx.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for k in range(K):
        step(k)      # every lane has drained when a step returns (evaluate_batch collects all results)
    if world > 1:     # the path's one exchange step: every rank's (1+P) doubles per evaluation, once
        dist.all_gather_into_tensor(gathered.view(-1), torch.from_numpy(results.reshape(-1)).to("cuda", non_blocking=True))
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = sum(cx.launch_count() for cx in ctxs)
    exec_flops = sum(cx.flop_count() for cx in ctxs) / (K * B)
    clocks = sampler.summary()
    if world > 1:
        tmax = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
    ms_step = ms_total / K
    ms_eval = ms_step / B
    value = world * K * B / (ms_total * 1e-3)
    out = results[K - 1, B - 1].copy()

    # ---- e2e: the public API (drop-in class) with HOST buffers, copies inside the timed region -----------------
    with contextlib.redirect_stdout(sys.stderr):      # the drop-in prints like the reference; stdout carries ONE JSON line
        model = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=train, test_nodes=test, theta=list(THETA), relabel_nodes=True,
                                          kerneltype=KERNEL, training_values=False)
    tv = pd.Series(t[train], index=train)
    model.set_training_values(tv)
    from gpnet_b200.engine import register_context
    register_context(ctx)             # same device context (same stream); the engine re-binds the graph in the warm-up calls
    model._engine.device = local
    for cx in ctxs[1:]:
        cx.close()                    # the engine owns its own extra lanes
    ctxs = [ctx]
    model.logPosterior_and_grad_batch(thetas[: 2 * J], train, tv, lanes=J)
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        vals, grads = model.logPosterior_and_grad_batch(thetas, train, tv, lanes=J)      # host thetas / targets in, host arrays out
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    val = float(vals[B - 1])
    if world > 1:
        tmax = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e_value = world * K * B / e2e_s
    health["e2e_vs_device_rel_diff"] = float(abs(val - out[0]) / abs(val))
    lane_ctxs = model._engine.lanes(J, model._positions(train), model._positions(test))

    roof = cpu = routes = None
    extras = {}
    if rank == 0:
        # ---- roofline of the dominant kernel: per-launch CUDA events over one extra (untimed) evaluation.  Launches of one
        #      kernel that run concurrently (one factorisation per part of the dissection, each on its own stream) share the chip:
        #      achieved = executed flops of all launches / length of the union of their intervals ----
        n_prof = 4 * J
        for cx in lane_ctxs:
            cx.set_targets(t[train])
            cx.set_profile(True)
        evaluate_batch(lane_ctxs, kind, thetas[:n_prof], None, True)
        iv = np.concatenate([cx.profile_intervals(ref=lane_ctxs[0]) for cx in lane_ctxs])
        for cx in lane_ctxs:
            cx.set_profile(False)
            cx.set_sm_budget(0)

        def union_ms(rows):
            total, cur0, cur1 = 0.0, 0.0, -1.0
            for a0, a1 in sorted((float(r[0]), float(r[1])) for r in rows):
                if cur1 < cur0 or a0 > cur1:
                    if cur1 >= cur0:
                        total += cur1 - cur0
                    cur0, cur1 = a0, a1
                elif a1 > cur1:
                    cur1 = a1
            return total + (cur1 - cur0 if cur1 >= cur0 else 0.0)
        algo = algorithmic_flops(N, n)
        fam = {}
        for name, kid in (("potrf_dataflow_kernel (persistent tile-dataflow Cholesky with appended solve / inverse rows, TMA + DMMA.8x8x4)", 2),
                          ("gemm_f64_kernel (TMA + DMMA.8x8x4, 128x128x16 / 64x64x16 tiles)", 0)):
            rows = iv[iv[:, 3] == kid]
            fam[name] = (union_ms(rows), float(rows[:, 2].sum()), len(rows) / n_prof, float((rows[:, 1] - rows[:, 0]).sum()))
        batch_ms = float(iv[:, 1].max() - iv[:, 0].min())
        kname = max(fam, key=lambda k: fam[k][0])
        k_un, k_fl, k_cnt, k_sum = fam[kname]
        k_tf = k_fl / (k_un * 1e-3) * 1e-12 if k_un > 0 else 0.0
        traffic, traffic_src = recorded_traffic(kname)
        roof = {"bound": "tensor", "kernel": kname,
                "achieved": k_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": k_tf / peak_tf if peak_tf else None,
                "peak_source": "measured in this run: register-resident DMMA.8x8x4 loop (gpn_fp64_peak); MEASURED_PEAKS.json has no "
                               "FP64 entry; nominal B200 FP64 is 40 TFLOP/s (37.2 at 148 SM x 64 FMA/clk x 1.965 GHz)",
                "how": f"one extra (untimed) batch of {n_prof} evaluations on the {J} lanes with a CUDA-event pair around every launch: executed "
                       "flops of all launches of the kernel / length of the union of their launch intervals (the launches of one "
                       "evaluation and of the lanes run concurrently on disjoint SM groups and share the chip)",
                "launches_per_eval": k_cnt, "union_ms": k_un, "sum_of_launch_ms": k_sum, "profiled_batch_ms": batch_ms,
                "flops_per_launch_avg": k_fl / max(k_cnt * n_prof, 1),
                "kernel_share_of_eval": k_un / batch_ms if batch_ms > 0 else None,
                "bound_note": "latency: every launch is bound by the per-tile-column dependency chain of a small factorisation "
                              "(DESIGN.md section 4), not by the tensor pipe",
                "other_kernels": {k: {"union_ms": v[0], "tflops": v[1] / (v[0] * 1e-3) * 1e-12 if v[0] > 0 else None,
                                      "launches_per_eval": v[2], "share_of_eval": v[0] / batch_ms if batch_ms > 0 else None}
                                  for k, v in fam.items() if k != kname},
                # utilisation is claimed from EXECUTED flops only (SURVEY 8d rule)
                "eval_executed_tflops": exec_flops / (ms_eval * 1e-3) * 1e-12,
                "eval_executed_frac_of_measured_peak": exec_flops / (ms_eval * 1e-3) * 1e-12 / peak_tf if peak_tf else None,
                "executed_gflop_per_eval": exec_flops * 1e-9,
                "conventional_gflop_per_eval": algo * 1e-9,
                "eval_conventional_tflops": algo / (ms_eval * 1e-3) * 1e-12,
                "traffic": traffic, "traffic_source": traffic_src,
                "route": {"level": route, "parts": parts}}
        if world == 1 and not args.no_extras:
            # the SAME two kernels at flop-bound sizes, measured live: what the kernels reach when the work, not a dependency chain
            # of a small factorisation, is the bound (M = I + 0.6 L~ of the C3 graph: a plain 8192 Cholesky; A A^T with lower tiles +
            # mirrored store: one Pade product of the N = 8192 expm)
            Nf = N
            ldf = (Nf + 127) // 128 * 128
            A_t = torch.zeros((ldf, ldf), dtype=torch.float64, device="cuda")
            C_t = torch.zeros((ldf, ldf), dtype=torch.float64, device="cuda")
            fb = {}
            for what in ("potrf", "pade_product"):
                best_ms = None
                for rep in range(3):
                    ctx.laplacian_dense(0.6 if what == "potrf" else -0.6, 1.0 if what == "potrf" else 0.0, A_t.data_ptr(), ldf)
                    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    r0.record(stream)
                    if what == "potrf":
                        info_f = ctx.potrf(Nf, A_t.data_ptr(), ldf)
                    else:
                        ctx.syrk(Nf, Nf, 1.0, A_t.data_ptr(), ldf, 0.0, C_t.data_ptr(), ldf)
                    r1.record(stream)
                    torch.cuda.synchronize()
                    ms = r0.elapsed_time(r1)
                    best_ms = ms if best_ms is None else min(best_ms, ms)
                T128 = (Nf + 127) // 128
                fl = Nf ** 3 / 3.0 if what == "potrf" else 2.0 * 128 * 128 * Nf * (T128 * (T128 + 1) // 2)
                fb[what] = {"n": Nf, "ms": best_ms, "tflops": fl / (best_ms * 1e-3) * 1e-12,
                            "frac_of_measured_peak": fl / (best_ms * 1e-3) * 1e-12 / peak_tf if peak_tf else None,
                            "flops": "n^3/3 (standard count)" if what == "potrf" else "executed: 2080 lower 128x128 tiles x K"}
            roof["same_kernels_at_flop_bound_sizes"] = {
                "potrf_dataflow_kernel, plain Cholesky": fb["potrf"], "gemm_f64_kernel, symmetric N x N x N product": fb["pade_product"],
                "note": "best of 3, CUDA events on the launching stream; the timed step's launches are 1200-1300-row factorisations"}
            del A_t, C_t
            torch.cuda.empty_cache()
        if world == 1:
            # the same evaluation through the other exact routes (same outputs to rounding), device-timed like `value`
            routes = {}
            names = {0: "full_kernel (K formed, gathered, K_y factored: the reference's order of operations)",
                     2: "schur (dense N x N Cholesky + triangular inverse)",
                     4: "schur_sparse_coupling_augmented (one factorisation of the other-node block)",
                     5: "schur_sparse_coupling_augmented_dissected (other nodes in two halves; round-1 default)",
                     6: "whole_graph_dissection (default: no dense Schur complement, one augmented factorisation per part), single lane"}
            for level in (0, 2, 4, 5, 6):
                ctx.set_fast_paths(level)
                for _ in range(2):
                    o2 = ctx.gpr_logp_grad(kind, thetas[B - 1], None, True)[1]
                ctx.reset_counters()
                kk = 8
                ctx.set_profile(True)
                ctx.gpr_logp_grad(kind, thetas[B - 1], None, True)
                pr = ctx.profile_summary()
                ctx.set_profile(False)
                ctx.reset_counters()
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record(stream)
                for _ in range(kk):
                    ctx.gpr_logp_grad(kind, thetas[B - 1], None, True)
                r1.record(stream)
                torch.cuda.synchronize()
                ms = r0.elapsed_time(r1) / kk
                fl = ctx.flop_count() / kk
                routes[names[level]] = {"value": 1e3 / ms, "ms_per_eval": ms, "executed_gflop_per_eval": fl * 1e-9,
                                        "executed_tflops": fl / (ms * 1e-3) * 1e-12,
                                        "frac_of_measured_peak": fl / (ms * 1e-3) * 1e-12 / peak_tf if peak_tf else None,
                                        "potrf_tflops_over_union": pr["potrf_flops"] / (pr["potrf_union_ms"] * 1e-3) * 1e-12 if pr["potrf_union_ms"] > 0 else None,
                                        "max_rel_diff_vs_default": float(np.max(np.abs(o2 - out) / np.abs(out)))}
            ctx.set_fast_paths(True)
            routes[names[6]]["note"] = "one evaluation at a time on the whole chip (latency); the headline runs %d of them at once" % J
            if not args.no_extras:
                # driver-visible expm evidence: one diffusion logp+grad at N = 8192 (7 N^3 executed: all products symmetric)
                th_d = np.log([0.6, 0.01])
                kd = KERNEL_IDS["diffusion"]
                ctx.gpr_logp_grad(kd, th_d, None, True)
                ctx.reset_counters()
                ctx.set_profile(True)
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record(stream)
                st, od, info = ctx.gpr_logp_grad(kd, th_d, None, True)
                r1.record(stream)
                torch.cuda.synchronize()
                pr = ctx.profile_summary()
                ctx.set_profile(False)
                ms = r0.elapsed_time(r1)
                fl = ctx.flop_count()
                m_used, s_used = ctx.expm_info()
                extras["diffusion_n8192_logp_grad"] = {
                    "ms": ms, "evals_per_s": 1e3 / ms, "pade_order": m_used, "squarings": s_used, "executed_gflop": fl * 1e-9,
                    "executed_tflops": fl / (ms * 1e-3) * 1e-12, "frac_of_measured_peak": fl / (ms * 1e-3) * 1e-12 / peak_tf if peak_tf else None,
                    "gemm_tflops_while_active": pr["gemm_flops"] / (pr["gemm_union_ms"] * 1e-3) * 1e-12 if pr["gemm_union_ms"] > 0 else None,
                    "gemm_share": pr["gemm_union_ms"] / ms, "neg_logp": float(od[0]),
                    "what": "GPnetRegressor diffusion kernel (expm by Pade + squaring, GPnet.py:338-340) logp+grad on the C3 graph, from scratch"}
        if world == 1 and not args.no_cpu_baseline:
            cores = pin_blas_threads()
            dt, cval = cpu_port_eval(problem, N)
            st, o0, _ = ctx.gpr_logp_grad(kind, THETA, None, True)
            cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"1 full-size logp+grad evaluation (N={N}, n={n}) of the dense numpy/scipy restatement with the analytic "
                             f"gradient ({dt:.1f} s); GPU/CPU -logp agree to {abs(cval - o0[0]) / abs(cval):.1e}"}
    if not args.no_extras:
        # C4 (the sharded landscape north_star names) rides along at every N so that the driver's scaling runs carry it
        reps = 2 if world == 1 else 3
        sec, lml, ev_ms = run_c4(ctx, stream, world, rank, local, reps, barrier)
        if rank == 0:
            extras["c4_landscape"] = {"seconds_per_landscape": sec, "device_event_ms": ev_ms, "points": int(lml.size),
                                      "points_per_s": lml.size / sec, "n_gpus": world, "scaling": "strong",
                                      "lml_max": float(np.max(lml[np.isfinite(lml)])), "argmax": [int(x) for x in np.unravel_index(np.argmax(lml), lml.shape)],
                                      "what": "GPnetRegressor.lml_landscape (GPnet.py:539-552), diffusion, RGG N=4096, n=2048, 64x64 grid over "
                                              "(theta0, noise): host axes in, host array out; kernel groups sharded over the ranks (longest "
                                              "first), one NCCL all-gather of (1+P) doubles per point; best of %d repetitions, max over ranks" % reps}
        # the same landscape on the spectral path: one eigendecomposition of L~ per graph, then one symmetric
        # product per kernel group
        first = []
        sec_s, lml_s, ev_s = run_c4(ctx, stream, world, rank, local, reps, barrier, spectral=True, first_out=first)
        if rank == 0:
            fin = np.isfinite(lml) & np.isfinite(lml_s)
            extras["c4_landscape_spectral"] = {"seconds_per_landscape": sec_s, "first_landscape_seconds_including_eigendecomposition": first[0],
                                               "jacobi_sweeps_and_offdiag_rel": list(first[1]) if first[1] else None,
                                               "points_per_s": lml_s.size / sec_s, "n_gpus": world,
                                               "max_rel_diff_vs_default_path": float(np.max(np.abs(lml_s[fin] - lml[fin]) / np.abs(lml[fin]))),
                                               "what": "use_spectral(): K = V exp(-beta lam) V^T from one GPU eigendecomposition (two-sided block "
                                                       "Jacobi) of the normalized Laplacian instead of one Pade expm per kernel group"}
        if rank == 0 and world == 1:
            # config C1 (30 x 30 grid, 90 training nodes): the launch-bound regime.  64 x 64 landscapes of the regressor and of the
            # classifier through the drop-in classes on the spectral path: ONE launch, one CTA per grid point
            try:
                extras["c1_batched_landscapes"] = c1_batched_landscapes(ctx, local)
            except Exception as e:      # evidence only: never lose the headline line over it
                extras["c1_batched_landscapes"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config("c3", N, n, P, world, B),
                "ms_per_eval": ms_eval, "timed_region_s": ms_total * 1e-3,
                "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": J * 8 * n + B * 8 * P,
                                          "d2h_bytes_per_step": B * (8 * 32 + 4 * 16),
                                          "api": "GPnetRegressor.logPosterior_and_grad_batch(thetas, nodes, Series, lanes=%d)" % J},
                "lanes": J, "health": health,
                "setup": {"first_evaluation_ms": first_eval_ms, "second_evaluation_ms": second_eval_ms,
                          "what": "wall clock of the first two single evaluations on lane 0: the first includes the host-side dissection "
                                  "analysis of the (graph, node set) pair and all buffer allocations"},
                "gpu_launches": int(launches), "roofline": roof}
        if routes is not None:
            line["routes"] = routes
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extras:
            line["extras"] = extras
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main_c4(args, ctx, stream, world, rank, local, W, K, barrier):
    """--workload c4: one step = one whole 64 x 64 landscape (strong scaling)."""
    import torch
    import torch.distributed as dist
    sampler = ClockSampler(local)
    run_c4(ctx, stream, world, rank, local, max(W - 1, 1), barrier)          # warm-up landscapes
    peak_tf = ctx.fp64_peak(0.3) if rank == 0 else 0.0
    sampler.start()
    ctx.reset_counters()
    secs = []
    for _ in range(K):
        sec, lml, ev = run_c4(ctx, stream, world, rank, local, 1, barrier)
        secs.append(sec)
    clocks = sampler.summary()
    launches = ctx.launch_count()
    flops = ctx.flop_count()
    if rank == 0:
        sec = float(np.mean(secs))
        # run_c4 runs one warm-up repetition per call: counters cover 2K landscapes on this rank
        line = {"metric": "logp landscape points/sec, 64x64 grid at N=4096 nodes (fp64)", "value": lml.size / sec, "unit": "points/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config("c4", 4096, 2048, 2, world, 1),
                "clocks": clocks, "e2e": {"value": lml.size / sec, "unit": "points/s", "h2d_bytes_per_step": 8 * (2048 + 2 + 128),
                                          "d2h_bytes_per_step": 8 * 3 * lml.size // world},
                "gpu_launches": int(launches // 2),
                "roofline": {"bound": "tensor", "kernel": "gemm_f64_kernel (Pade products of expm)", "achieved": flops / 2 / K / sec * 1e-12,
                             "peak": peak_tf, "unit": "TFLOP/s", "frac": flops / 2 / K / sec * 1e-12 / peak_tf if peak_tf else None,
                             "how": "executed flops of this rank's share of one landscape / seconds per landscape", "traffic": None}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
