"""GPU tier: the spectral fast path (K = V r(lam) V^T from one eigendecomposition of the normalized Laplacian)
and the batched small-n evaluator (8f-3: one CTA per theta) against the numpy oracle and against the default CUDA path.
Own parity budget: the same 1e-9 norm-wise bound as the default path (measured ~1e-12)."""
import os
import time

import numpy as np
import pandas as pd
import pytest

from conftest import golden_files, load_golden, nrel
from oracle import configs as cfg
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu
TOL = 1e-9

KERNELS = [("diffusion", np.log([0.6, 0.01])), ("regularized_laplacian", np.log([0.6, 0.01])), ("pstep_walk", np.log([2.3, 4, 0.01]))]


def regressor(G, tr, te, t, kt, th, spectral):
    import gpnet_b200
    m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=tr, test_nodes=te, theta=list(th), relabel_nodes=True, kerneltype=kt,
                                  training_values=False)
    m.set_training_values(pd.Series(t[tr], index=tr))
    if spectral:
        m.use_spectral()
    return m


@pytest.mark.parametrize("case", ["grid30", "rgg700", "ba500", "two_components"])
def test_jacobi_eigendecomposition_of_the_normalized_laplacian(gpu_ctx, case):
    """L~ = V diag(lam) V^T to working precision, V orthogonal, lam equal to LAPACK's; degenerate spectra (grid), hubs (BA),
    several zero eigenvalues (disconnected graph), N not a multiple of 64 (padding)."""
    import networkx as nx
    from gpnet_b200.engine import graph_csr
    if case == "grid30":
        G = cfg.relabel(nx.grid_2d_graph(30, 30))
    elif case == "rgg700":
        G = cfg.rgg_case(700, 100, 100)[0]
    elif case == "ba500":
        G = nx.barabasi_albert_graph(500, 3, seed=0)
    else:
        G = nx.disjoint_union(cfg.relabel(nx.grid_2d_graph(9, 7)), nx.path_graph(40))
        G.add_node(len(G))       # an isolated node: zero row of L~
    ip, ix, w, _, _ = graph_csr(G)
    N = len(ip) - 1
    gpu_ctx.set_graph(ip, ix, w)
    sweeps, off = gpu_ctx.spectral_prepare()
    lam, VT = gpu_ctx.spectral_get()
    L = go.normalized_laplacian(ip, ix, w, N)
    print(case, "sweeps", sweeps, "offdiag_rel", off)
    assert off < 1e-12 and 0 < sweeps <= 40
    assert np.max(np.abs(VT.T @ (lam[:, None] * VT) - L)) < 2e-11      # ~15 sweeps of accumulated rotations
    assert np.max(np.abs(VT @ VT.T - np.eye(N))) < 2e-11
    assert np.max(np.abs(lam - np.linalg.eigvalsh(L))) < 2e-11
    gpu_ctx.spectral_enable(False)


@pytest.mark.parametrize("kt,th", KERNELS)
def test_spectral_kernel_blocks_match_the_oracle(kt, th):
    G, tr, te, t = cfg.rgg_case(700, 300, 200)
    m = regressor(G, tr, te, t, kt, th, True)
    orc = go.OracleGP(G, kerneltype=kt, expm_mode="dense")
    K = orc.full(th)
    k = m.kernel(tr, tr, th)
    ks = m.kernel(te, tr, th, measnoise=False)
    c = np.exp(th[-1])
    assert nrel(k, K[np.ix_(tr, tr)] + c) <= TOL
    assert nrel(ks, K[np.ix_(te, tr)]) <= TOL
    assert m._engine.ctx.expm_info() == (0, 0)        # no Pade product ran: the kernel came from the eigendecomposition


@pytest.mark.parametrize("path", [p for p in golden_files("reg_") if "grid15" in p or "rgg1024" in p] or golden_files("reg_")[:3], ids=os.path.basename)
def test_spectral_regressor_against_the_reference_goldens(path):
    """predict / logp / gradient of the drop-in regressor with the spectral kernels against the shimmed reference's outputs."""
    from test_gpu_parity import make_regressor
    g = load_golden(path)
    m = make_regressor(g).use_spectral()
    assert m.predict() is m and m.is_trained
    assert nrel(m.fstar, g["fstar"]) <= TOL and nrel(m.s, g["s"]) <= TOL and nrel(m.alpha, g["alpha"]) <= TOL
    assert abs(m.logp() - g["logp"]) <= TOL * abs(g["logp"])
    if "neg_logp_grad_analytic_oracle" in g:
        grad = m.gradLogPosterior(m.theta, m.training_nodes, m.training_values)
        assert nrel(grad, g["neg_logp_grad_analytic_oracle"]) <= 1e-8


@pytest.mark.parametrize("path", [p for p in golden_files("cls_") if "ba400" in p] or golden_files("cls_")[:2], ids=os.path.basename)
def test_spectral_classifier_against_the_reference_goldens(path):
    import gpnet_b200
    from test_gpu_parity import graph_from_csr
    g = load_golden(path)
    tr, te = [int(x) for x in g["train"]], [int(x) for x in g["test"]]
    y = pd.Series(g["y_train"].astype(np.int64), index=tr)
    m = gpnet_b200.GPnetClassifier(Graph=graph_from_csr(g), training_nodes=tr, test_nodes=te, theta=list(g["theta"]), relabel_nodes=True,
                                   kerneltype=str(g["kerneltype"]), training_values=y).use_spectral()
    fstar, V, probs = m.predict()
    assert m.newton_iterations == g["iters"]
    assert nrel(fstar, g["fstar"]) <= TOL and nrel(V, g["V"]) <= TOL and nrel(probs, g["probs"]) <= TOL
    assert np.array_equal(np.asarray(m.df["predicted_class"][te], dtype=np.float64), g["labels"])


@pytest.mark.parametrize("kt,th", KERNELS)
def test_batched_small_n_evaluator_c1_against_the_oracle(kt, th):
    """Config C1 (30 x 30 grid, 90 training nodes): a batch of thetas in one launch (one CTA each) against the oracle's
    -logp and analytic gradient, point by point, and against the default CUDA path."""
    G, tr, te, t = cfg.grid_case(30, 30, 90, 810)
    m = regressor(G, tr, te, t, kt, th, True)
    rng = np.random.RandomState(3)
    thetas = [np.array(th, dtype=float)]
    for _ in range(5):
        d = rng.uniform(-0.3, 0.3, len(th))
        if kt == "diffusion":
            d[0] = -abs(d[0])              # beta < 1
        if kt == "pstep_walk":
            d[0] = abs(d[0])               # a >= 2
        thetas.append(np.array(th) + d)
    vals, grads = m.logPosterior_and_grad_batch(thetas, tr, m.training_values)
    assert m._engine.ctx.last_route()[0] == 7
    orc = go.OracleGP(G, kerneltype=kt, expm_mode="dense")
    ref = regressor(G, tr, te, t, kt, th, False)
    for b, thb in enumerate(thetas):
        oval, ograd = orc.neg_logp_grad(thb, tr, t[tr])
        assert abs(vals[b] - oval) <= TOL * abs(oval), (kt, b)
        assert nrel(grads[b], ograd) <= 1e-8, (kt, b)
        v2, g2 = ref.logPosterior_and_grad(thb, tr, ref.training_values)
        assert abs(vals[b] - v2) <= TOL * abs(v2) and nrel(grads[b], g2) <= 1e-8


def test_batched_evaluator_reports_a_failed_cholesky_and_domain_errors():
    G, tr, te, t = cfg.grid_case(12, 12, 50, 60)
    m = regressor(G, tr, te, t, "regularized_laplacian", np.log([0.6, 0.01]), True)
    vals, grads = m.logPosterior_and_grad_batch([np.log([0.6, 0.01]), np.array([np.log(0.6), -800.0])], tr, m.training_values)
    assert np.isfinite(vals[0])                      # exp(-800) = 0 noise: still PD for this kernel
    m2 = regressor(G, tr, te, t, "diffusion", np.log([0.6, 0.01]), True)
    with pytest.raises(AssertionError):
        m2.logPosterior_and_grad_batch([np.log([1.5, 0.01])], tr, m2.training_values)


def test_c1_landscape_through_the_batched_evaluator_and_amortised_time():
    """The 64 x 64 landscape of C1 over (theta0, noise) in one launch equals the default path on a sub-grid, and one
    evaluation costs < 150 us amortised (the default path takes 0.3 - 1.5 ms per evaluation)."""
    import torch
    G, tr, te, t = cfg.grid_case(30, 30, 90, 810)
    th = np.log([0.6, 0.01])
    ax1 = np.linspace(np.log(0.01), np.log(0.99), 64)
    ax2 = np.linspace(np.log(0.001), np.log(10.0), 64)
    m = regressor(G, tr, te, t, "diffusion", th, True)
    lml = m.lml_landscape(list(th), [0, 1], ax1, ax2)               # includes the eigendecomposition
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lml = m.lml_landscape(list(th), [0, 1], ax1, ax2)
    dt = time.perf_counter() - t0
    assert lml.shape == (64, 64) and np.all(np.isfinite(lml))
    ref = regressor(G, tr, te, t, "diffusion", th, False)
    sub = ref.lml_landscape(list(th), [0, 1], ax1[::16], ax2[::8])
    assert nrel(lml[::16, ::8], sub) <= TOL
    orc = go.OracleGP(G, kerneltype="diffusion", expm_mode="dense")
    for (i, j) in [(0, 0), (63, 63), (40, 7)]:
        assert abs(lml[i, j] + orc.neg_logp(np.array([ax1[i], ax2[j]]), tr, t[tr])) <= TOL * abs(lml[i, j])
    per_eval_us = dt / lml.size * 1e6
    print("C1 landscape: %.1f ms for 4096 points = %.2f us per logp, Jacobi: %s" % (dt * 1e3, per_eval_us, m._engine.spectral_info))
    assert per_eval_us < 150.0


def test_spectral_landscape_n2048_equals_default_path():
    """Kernel groups of a landscape on a mid-size graph (n > 112: spectral kernel build + the usual Cholesky route)."""
    G, tr, te, t = cfg.rgg_case(2048, 1024, 0)
    th = np.log([0.6, 0.01])
    ax1 = np.linspace(np.log(0.05), np.log(0.9), 3)
    ax2 = np.linspace(np.log(0.001), np.log(1.0), 4)
    a = regressor(G, tr, [], t, "diffusion", th, True).lml_landscape(list(th), [0, 1], ax1, ax2)
    b = regressor(G, tr, [], t, "diffusion", th, False).lml_landscape(list(th), [0, 1], ax1, ax2)
    assert nrel(a, b) <= TOL


@pytest.mark.parametrize("N,n", [(70, 7), (130, 1), (333, 111), (400, 112)])
def test_batched_evaluator_ragged_sizes_and_weighted_graph(N, n):
    """Training-set sizes that are not multiples of the 4 x 4 register tiles (down to one node, up to the 112-node limit), a graph
    size that needs padding in the eigensolver, and edge weights (GPnet.py:329 honours the "weight" attribute)."""
    import networkx as nx
    rng = np.random.RandomState(N)
    G = nx.connected_watts_strogatz_graph(N, 6, 0.3, seed=1)
    for u, v in G.edges():
        G[u][v]["weight"] = float(rng.uniform(0.5, 2.0))
    tr, te = cfg.split(N, n, min(20, N - n))
    t = rng.randn(N)
    for kt, th in KERNELS:
        m = regressor(G, tr, te, t, kt, th, True)
        thetas = [np.array(th, dtype=float), np.array(th, dtype=float) + 0.1 * np.array([(-1.0 if kt == "diffusion" else 1.0)] + [0.0] * (len(th) - 2) + [0.5])]
        vals, grads = m.logPosterior_and_grad_batch(thetas, tr, m.training_values)
        assert m._engine.ctx.last_route()[0] == 7
        orc = go.OracleGP(G, kerneltype=kt, expm_mode="dense")
        for b, thb in enumerate(thetas):
            oval, ograd = orc.neg_logp_grad(thb, tr, t[tr])
            assert abs(vals[b] - oval) <= TOL * max(abs(oval), 1.0), (kt, N, n, b)
            assert np.max(np.abs(grads[b] - ograd)) <= 1e-8 * max(np.max(np.abs(ograd)), 1.0), (kt, N, n, b)


def test_training_sets_above_the_batch_limit_fall_back_to_the_lanes():
    G, tr, te, t = cfg.grid_case(20, 20, 150, 100)
    th = np.log([0.6, 0.01])
    m = regressor(G, tr, te, t, "regularized_laplacian", th, True)
    vals, grads = m.logPosterior_and_grad_batch([th, th + 0.05], tr, m.training_values, lanes=2)
    assert m._engine.ctx.last_route()[0] != 7
    orc = go.OracleGP(G, kerneltype="regularized_laplacian", expm_mode="dense")
    for b, thb in enumerate([th, th + 0.05]):
        oval, ograd = orc.neg_logp_grad(thb, tr, t[tr])
        assert abs(vals[b] - oval) <= TOL * abs(oval) and nrel(grads[b], ograd) <= 1e-8


@pytest.mark.parametrize("kt,th", [("diffusion", np.log([0.6, 0.1])), ("regularized_laplacian", np.log([0.6, 0.1])), ("pstep_walk", np.log([2.3, 4, 0.1]))])
def test_batched_newton_loop_against_the_oracle_and_the_default_path(kt, th):
    """GPnetClassifier.NRiteration for a batch of thetas in one launch (one CTA each): logq to 1e-9 and IDENTICAL Newton iteration
    counts against the oracle's loop and against the default CUDA loop."""
    import gpnet_b200
    from gpnet_b200._lib import KERNEL_IDS
    G, tr, te, y = cfg.ba_case(300, 100, 100)
    ys = pd.Series(y[tr], index=tr)
    m = gpnet_b200.GPnetClassifier(Graph=G, training_nodes=tr, test_nodes=te, theta=list(th), relabel_nodes=True, kerneltype=kt,
                                   training_values=ys).use_spectral()
    rng = np.random.RandomState(5)
    thetas = [np.array(th, dtype=float)]
    for _ in range(4):
        d = rng.uniform(-0.4, 0.4, len(th))
        if kt == "diffusion":
            d[0] = -abs(d[0])
        if kt == "pstep_walk":
            d[0] = abs(d[0])
        thetas.append(np.array(th) + d)
    ctx = m._engine.bind(m._positions(tr), m._positions(te))
    st, logq, iters, infos = ctx.gpc_newton_batch(KERNEL_IDS[kt], np.asarray(thetas), y[tr].astype(float))
    assert st == 0 and np.all(infos == 0)
    orc = go.OracleGP(G, kerneltype=kt, expm_mode="dense")
    ref = gpnet_b200.GPnetClassifier(Graph=G, training_nodes=tr, test_nodes=te, theta=list(th), relabel_nodes=True, kerneltype=kt,
                                     training_values=ys)
    for b, thb in enumerate(thetas):
        o_f, o_logq, o_a, o_iters = orc.newton(thb, tr, y[tr])[:4]
        assert abs(logq[b] - o_logq) <= TOL * abs(o_logq), (kt, b, logq[b], o_logq)
        assert int(iters[b]) == o_iters, (kt, b)
        f, lq, a = ref.NRiteration(tr, ys, thb)
        assert abs(logq[b] - float(lq[0, 0])) <= TOL * abs(logq[b]) and int(iters[b]) == ref.newton_iterations


def test_classifier_landscape_through_the_batched_newton_loop():
    import gpnet_b200
    G, tr, te, y = cfg.ba_case(300, 100, 100)
    ys = pd.Series(y[tr], index=tr)
    th = np.log([0.6, 0.1])
    ax1 = np.linspace(np.log(0.05), np.log(0.9), 6)
    ax2 = np.linspace(np.log(0.01), np.log(1.0), 5)

    def model(spectral):
        m = gpnet_b200.GPnetClassifier(Graph=G, training_nodes=tr, test_nodes=te, theta=list(th), relabel_nodes=True, kerneltype="diffusion",
                                       training_values=ys)
        return m.use_spectral() if spectral else m
    a = model(True).lml_landscape(list(th), [0, 1], ax1, ax2)
    b = model(False).lml_landscape(list(th), [0, 1], ax1, ax2)
    assert a.shape == (6, 5) and nrel(a, b) <= TOL
