"""Headline benchmark: GP logp+grad evaluations per second at N = 8192 graph nodes (fp64), config C3 of SURVEY.md
Appendix D (GPnetRegressor, regularized-Laplacian kernel, random geometric graph, n = 4096 training nodes, P = 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nodes 8192]

One "step" = one evaluation theta (host) -> (-logp, -dlogp/dtheta) (host), from scratch: nothing is carried over between
evaluations (like the reference, which rebuilds everything per call); only the graph, the node sets and the targets are
resident.  The default route reads the inverse training block off a Schur complement instead of forming the kernel
(DESIGN.md section 4); `routes` in the JSON line times the reference's own order of operations (full kernel, block gather,
Cholesky, alpha, log-det, K_y^-1, the P trace terms) and the intermediate routes beside it, and utilisation is always
reported from EXECUTED flops.  Prints ONE JSON line (see the task contract); rank 0 only.
"""
import argparse
import json
import math
import os
import subprocess
import sys
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


def workload_config(N, n, P, world):
    """The same `config` object on both arms."""
    return {"workload": f"C3: GPnetRegressor regularized_laplacian, random geometric graph N={N}, n={n}, P={P}, theta=log[0.6,0.01]; "
                        f"one logp+grad per step per GPU (multi-start thetas across GPUs, (1+P)-double NCCL all-gather per step)",
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


def cpu_port_eval(problem, N, reference_style):
    """CPU legs (the ONLY place bench.py executes oracle/).  reference_style=True: one logPosterior the way the
    reference computes it (scipy.linalg.inv + cholesky + two general solves, GPnet.py:342 + GPnetRegressor.py:139-150);
    False: the best-effort dense restatement with the analytic gradient (one full logp+grad)."""
    from oracle import gp_oracle as go
    _, csr, train, test, t = problem
    orc = go.OracleGP(csr=csr, kerneltype=KERNEL, expm_mode="dense")
    tt = t[train]
    t0 = time.perf_counter()
    if reference_style:
        val = orc.neg_logp(THETA, train, tt)
    else:
        val, _ = orc.neg_logp_grad(THETA, train, tt)
    return time.perf_counter() - t0, val


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([d.get("num_threads", 1) for d in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count()


def run_reference(args):
    """--impl reference: the reference's own CPU path (the oracle port with the reference's call pattern; the pure-Python
    reference cannot travel to the GPU box) on all host cores.  The reference's optimiser gets gradients by finite
    differences: (P+1) logPosterior calls per logp+grad (GPnet.py:264-271 passes no jac).  Each step times ONE such
    logPosterior probe at full size; value = 1 / ((P+1) * t_step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N = args.nodes
    problem = c3_problem(N)
    P = len(THETA)
    budget = 200.0
    times = []
    t_begin = time.perf_counter()
    for i in range(args.warmup + args.steps):
        dt, _ = cpu_port_eval(problem, N, reference_style=True)
        if i >= args.warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > budget and len(times) >= 1:
            break
    t_step = float(np.mean(times))
    value = 1.0 / ((P + 1) * t_step)
    cores = blas_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": args.warmup, "ms_per_step": t_step * 1e3 * (P + 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(N, N // 2, P, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{len(times)} x one full-size logPosterior probe (N={N}); a logp+grad costs (P+1)=3 probes "
                                       "by finite differences as in the reference's optimize_params"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--nodes", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    N = args.nodes
    n = N // 2
    problem = c3_problem(N)
    G, csr, train, test, t = problem
    P = len(THETA)
    kind = KERNEL_IDS[KERNEL]

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = Context(local, stream.cuda_stream)
    ctx.set_graph(*csr)
    ctx.set_nodes(train, test)
    ctx.set_targets(t[train])
    # multi-start: every rank evaluates its own theta (weak scaling: per-GPU work fixed), results all-gathered
    theta = THETA + 0.01 * rank * np.array([1.0, -1.0])
    gathered = torch.zeros((world, 1 + P), dtype=torch.float64, device="cuda") if world > 1 else None

    def step():
        st, out, info = ctx.gpr_logp_grad(kind, theta, None, True)
        assert st == 0 and info == 0, (st, info)
        if world > 1:     # the path's real exchange step: (1+P) doubles per rank
            dist.all_gather_into_tensor(gathered, torch.from_numpy(out).to("cuda", non_blocking=True).reshape(1, -1))
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        out = step()
    barrier()
    peak_tf = ctx.fp64_peak(0.3) if rank == 0 else 0.0
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(K):
        out = step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count()
    exec_flops = ctx.flop_count() / K
    clocks = sampler.summary()
    if world > 1:
        tmax = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
    ms_step = ms_total / K
    value = world * K / (ms_total * 1e-3)

    # ---- e2e: the public API (drop-in class) with HOST buffers, copies inside the timed region -----------------
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):      # the drop-in prints like the reference; stdout carries ONE JSON line
        model = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=train, test_nodes=test, theta=list(theta), relabel_nodes=True,
                                          kerneltype=KERNEL, training_values=False)
    tv = pd.Series(t[train], index=train)
    model.set_training_values(tv)
    from gpnet_b200.engine import register_context
    register_context(ctx)             # same device context (same stream); the engine re-binds the graph in the warm-up calls
    model._engine.device = local
    for _ in range(2):
        model.logPosterior_and_grad(theta, train, tv)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        val, grad = model.logPosterior_and_grad(theta, train, tv)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e_value = world * K / e2e_s
    assert abs(val - out[0]) <= 1e-12 * abs(val)

    # ---- roofline of the dominant kernel: per-launch CUDA events on the context's stream over one extra (untimed) step ----
    roof = None
    cpu = None
    routes = None
    if rank == 0:
        ctx.set_profile(True)
        ctx.gpr_logp_grad(kind, theta, None, True)
        prof = ctx.profile_summary()
        ctx.set_profile(False)
        algo = algorithmic_flops(N, n)
        cands = {
            "gemm_f64_kernel<2,4,6,*,*> (TMA + DMMA.8x8x4, 128x128x16 tiles)": (prof["big_ms"], prof["big_flops"], prof["big_launches"]),
            "potrf_dataflow_kernel (persistent tile-dataflow Cholesky, TMA + DMMA.8x8x4)": (prof["potrf_ms"], prof["potrf_flops"],
                                                                                             prof["potrf_launches"]),
        }
        kname = max(cands, key=lambda k: cands[k][0])
        k_ms, k_fl, k_cnt = cands[kname]
        k_tf = k_fl / (k_ms * 1e-3) * 1e-12 if k_ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": kname,
                "achieved": k_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": k_tf / peak_tf if peak_tf else None,
                "peak_source": "measured in this run: register-resident DMMA.8x8x4 loop (gpn_fp64_peak); MEASURED_PEAKS.json has no "
                               "FP64 entry; nominal B200 FP64 is 40 TFLOP/s (37.2 at 148 SM x 64 FMA/clk x 1.965 GHz)",
                "flops_per_launch_avg": k_fl / max(k_cnt, 1), "launch_ms_avg": k_ms / max(k_cnt, 1), "launches_per_step": k_cnt,
                "kernel_share_of_step": k_ms / ms_step,
                "other_kernels": {k: {"ms_per_step": v[0], "tflops": v[1] / (v[0] * 1e-3) * 1e-12 if v[0] > 0 else None,
                                      "launches": v[2], "share_of_step": v[0] / ms_step} for k, v in cands.items() if k != kname},
                # utilisation is claimed from EXECUTED flops only (SURVEY 8d rule): the default route of this workload reads
                # K_tt^-1 off the Cholesky factor and executes about half of the conventional count below
                "step_executed_tflops": exec_flops / (ms_step * 1e-3) * 1e-12,
                "step_executed_frac_of_measured_peak": exec_flops / (ms_step * 1e-3) * 1e-12 / peak_tf if peak_tf else None,
                "step_executed_frac_of_nominal_40tf": exec_flops / (ms_step * 1e-3) * 1e-12 / 40.0,
                "executed_gflop_per_step": exec_flops * 1e-9,
                "conventional_gflop_per_step": algo * 1e-9,
                "step_conventional_tflops": algo / (ms_step * 1e-3) * 1e-12,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures of the two launches
                # of this kernel in one step (profiles/r01_ncu_potrf_half.txt: 0.019 GB for each half-block launch,
                # profiles/r01_ncu_potrf_tt.txt: 3.033 GB);
                # recorded, not re-measured here (numbers taken under a profiler are never bench values)
                "traffic": 1.024e9 if kname.startswith("potrf_dataflow") and N == 8192 else None,
                "traffic_unit": "bytes per launch (average of the step's three launches; ncu captures in profiles/)"}
        # the same evaluation through the two other exact routes (same outputs to rounding), device-timed like `value`
        if world == 1:
            routes = {}
            names = {0: "full_kernel (K formed, gathered, K_y factored: the reference's order of operations)",
                     1: "block (training nodes last, only K[tr,:] formed)", 2: "schur (dense N x N Cholesky + triangular inverse)",
                     3: "schur_sparse_coupling (separate triangular inverse / product launches)",
                     4: "schur_sparse_coupling_augmented (one factorisation of the other-node block)",
                     5: "schur_sparse_coupling_augmented_dissected (other nodes in two halves)",
                     6: "whole_graph_dissection (default: no dense Schur complement, one augmented factorisation per part)"}
            for level in (0, 1, 2, 3, 4, 5):
                ctx.set_fast_paths(level)
                for _ in range(2):
                    o2 = ctx.gpr_logp_grad(kind, theta, None, True)[1]
                ctx.reset_counters()
                kk = max(4, K // 4)
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record(stream)
                for _ in range(kk):
                    ctx.gpr_logp_grad(kind, theta, None, True)
                r1.record(stream)
                torch.cuda.synchronize()
                ms = r0.elapsed_time(r1) / kk
                routes[names[level]] = {"value": 1e3 / ms, "ms_per_step": ms, "executed_gflop_per_step": ctx.flop_count() / kk * 1e-9,
                                        "max_rel_diff_vs_default": float(np.max(np.abs(o2 - out) / np.abs(out)))}
            ctx.set_fast_paths(6)
            routes[names[6]] = {"value": value, "ms_per_step": ms_step, "executed_gflop_per_step": exec_flops * 1e-9}
        if world == 1 and not args.no_cpu_baseline:
            dt, cval = cpu_port_eval(problem, N, reference_style=False)
            cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": blas_threads(), "kind": "port",
                   "sample": f"1 full-size logp+grad evaluation (N={N}, n={n}) of the dense numpy/scipy restatement with the analytic "
                             f"gradient ({dt:.1f} s); GPU/CPU -logp agree to {abs(cval - out[0]) / abs(cval):.1e}"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(N, n, P, world),
                "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * (n + P) + 4 * 0,
                                          "d2h_bytes_per_step": 8 * 32 + 4 * 16},
                "gpu_launches": int(launches), "roofline": roof}
        if routes is not None:
            line["routes"] = routes
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
