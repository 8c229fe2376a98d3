"""GPU tier: parity of the CUDA path (through the drop-in classes / the C-ABI) against the golden vectors of
the shimmed reference and against the numpy oracle on the same seeded inputs.  Tolerance: 1e-9 norm-wise
relative on kernel entries, predictive mean / deviation and logp (north_star); classifier labels and Newton
iteration counts identical."""
import os

import networkx as nx
import numpy as np
import pandas as pd
import pytest

from conftest import golden_files, load_golden, nrel
from oracle import configs as cfg
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu
TOL = 1e-9


def graph_from_csr(g):
    N = len(g["indptr"]) - 1
    G = nx.Graph()
    G.add_nodes_from(range(N))
    for i in range(N):
        for e in range(g["indptr"][i], g["indptr"][i + 1]):
            j = int(g["indices"][e])
            if j >= i:
                G.add_edge(i, j, weight=float(g["weights"][e]))
    return G


def make_regressor(g):
    import gpnet_b200
    tr, te = [int(x) for x in g["train"]], [int(x) for x in g["test"]]
    m = gpnet_b200.GPnetRegressor(Graph=graph_from_csr(g), training_nodes=tr, test_nodes=te, theta=list(g["theta"]),
                                  relabel_nodes=True, kerneltype=str(g["kerneltype"]), training_values=False)
    m.set_training_values(pd.Series(g["t_train"], index=tr))
    return m


@pytest.mark.parametrize("path", golden_files("reg_"), ids=os.path.basename)
def test_regressor_parity(path):
    g = load_golden(path)
    m = make_regressor(g)
    assert m.predict() is m and m.is_trained
    assert nrel(m.fstar, g["fstar"]) <= TOL
    assert nrel(m.s, g["s"]) <= TOL
    assert nrel(m.alpha, g["alpha"]) <= TOL
    assert nrel(m.kstarstar_diag, g["kstarstar_diag"]) <= TOL
    assert list(m.df.columns) == ["pvtdist", "train_vals", "fstar", "variance_s", "prob_0", "prob_1", "predicted_class"]
    assert nrel(m.df["fstar"][m.test_nodes].values, g["fstar"]) <= TOL
    if "k" in g:        # lazily fetched device-resident attributes
        assert nrel(m.k, g["k"]) <= TOL and nrel(m.kstar, g["kstar"]) <= TOL
        L = m.L
        assert nrel(L @ L.T, g["k"]) <= TOL and np.all(np.triu(L, 1) == 0)
        v = m.v
        assert nrel(np.sum(v * v, axis=0), g["kstarstar_diag"] - g["s"] ** 2) <= 1e-8
        assert m.V.shape == (len(m.test_nodes),) * 2 and m.kstarstar.shape == m.V.shape
        # GPnetBase.kernel: fresh host array, graph order, noise on every entry
        kk = m.kernel(m.training_nodes, m.training_nodes, m.theta)
        assert nrel(kk, g["k"]) <= TOL
        ks = m.kernel(m.test_nodes, m.training_nodes, m.theta, measnoise=False)
        assert nrel(ks, g["kstar"]) <= TOL
    lp = m.logp()
    assert abs(lp - g["logp"]) <= TOL * abs(g["logp"])
    tv = m.training_values
    if "neg_logp_grad_fd" in g:
        grad = m.gradLogPosterior(m.theta, m.training_nodes, tv)
        sel = g["grad_fd_valid"]
        assert nrel(grad[sel], g["neg_logp_grad_fd"][sel]) <= 1e-6            # vs central differences of the reference
        assert nrel(grad, g["neg_logp_grad_analytic_oracle"]) <= 1e-8          # vs the analytic oracle
        val, grad2 = m.logPosterior_and_grad(m.theta, m.training_nodes, tv)
        assert abs(-val - g["logp"]) <= TOL * abs(g["logp"]) and nrel(grad2, grad) <= 1e-12
    if "land_lml" in g:
        th = list(g["theta"])
        lml = m.lml_landscape(th, list(g["land_axidx"]), g["land_ax1"], g["land_ax2"])
        assert nrel(lml, g["land_lml"]) <= TOL
        assert th[g["land_axidx"][0]] == g["land_ax1"][-1] and th[g["land_axidx"][1]] == g["land_ax2"][-1]   # A-14


@pytest.mark.parametrize("path", golden_files("cls_"), ids=os.path.basename)
def test_classifier_parity(path):
    import gpnet_b200
    g = load_golden(path)
    tr, te = [int(x) for x in g["train"]], [int(x) for x in g["test"]]
    y = pd.Series(g["y_train"].astype(np.int64), index=tr)
    m = gpnet_b200.GPnetClassifier(Graph=graph_from_csr(g), training_nodes=tr, test_nodes=te, theta=list(g["theta"]),
                                   relabel_nodes=True, kerneltype=str(g["kerneltype"]), training_values=y)
    f, logq, a = m.NRiteration(tr, y, g["theta"])
    assert f.shape == (len(tr), 1) and a.shape == (len(tr), 1) and np.shape(logq) == (1, 1)
    assert m.newton_iterations == g["iters"]
    assert abs(float(logq[0, 0]) - g["logq"]) <= TOL * abs(g["logq"])
    assert nrel(f.ravel(), g["f"]) <= TOL and nrel(a.ravel(), g["a"]) <= TOL
    fstar, V, probs = m.predict()
    assert m.newton_iterations == g["iters"]
    assert nrel(fstar, g["fstar"]) <= TOL and nrel(V, g["V"]) <= TOL and nrel(probs, g["probs"]) <= TOL
    labels = np.asarray(m.df["predicted_class"][te], dtype=np.float64)
    assert np.array_equal(labels, g["labels"])                   # identical classifier labels
    assert abs(float(np.squeeze(m.logp())) - g["logp"]) <= TOL * abs(g["logp"])


def test_demo_driver_flow_through_module_aliases(capsys):
    """A driver in the shape of the reference's graph_test.py (random node assignment from a seed, regressor on a lattice,
    labels from the pivot distance, classifier on the same split, every plot_* call) runs unchanged through the module
    aliases -- without matplotlib the plots print a note -- and its numbers match the oracle on the same split."""
    import gpnet_b200
    gpnet_b200.install_dropin()
    from GPnetRegressor import GPnetRegressor
    from GPnetClassifier import GPnetClassifier
    G = nx.generators.lattice.grid_graph(dim=[15, 15], periodic=False)
    theta = [np.log(2.3), np.log(4), np.log(0.01)]
    gpr = GPnetRegressor(Graph=G, ntrain=125, ntest=100, theta=theta, seed=123, kerneltype="pstep_walk", relabel_nodes=True)
    gpr.plot_graph()
    assert gpr.predict() is gpr
    gpr.plot_predict_2d()
    gpr.plot_predict_graph()
    gpr.plot_prior()
    gpr.plot_post()
    assert len(gpr.training_nodes) == 125 and len(gpr.test_nodes) == 100 and len(gpr.other_nodes) == 0
    tr, te = sorted(gpr.training_nodes), sorted(gpr.test_nodes)
    orc = go.OracleGP(gpr.Graph, kerneltype="pstep_walk")
    tvals = np.asarray([gpr.training_values[x] for x in tr], dtype=np.float64)
    ref = orc.predict_regressor(theta, tr, te, tvals)
    assert nrel(gpr.fstar, ref["fstar"]) <= TOL and nrel(gpr.s, ref["s"]) <= TOL

    labels = (np.sin(0.5 * gpr.pivot_distance(0)) > 0).replace({True: 1, False: -1})
    train_labels = labels[gpr.training_nodes]
    th_c = [np.log(2.1), np.log(2), np.log(0.1)]
    gpc = GPnetClassifier(Graph=G, training_nodes=gpr.training_nodes, test_nodes=gpr.test_nodes, training_values=train_labels,
                          theta=th_c, seed=321, kerneltype="pstep_walk", relabel_nodes=True)
    fstar, V, probs = gpc.predict()
    gpc.plot_graph()
    gpc.plot_latent()
    gpc.plot_predict_graph()
    gpc.plot_binary_prediction()
    y = np.asarray([train_labels[x] for x in tr], dtype=np.float64)
    refc = orc.predict_classifier(th_c, tr, te, y)
    assert gpc.newton_iterations == refc["iters"]
    assert nrel(fstar, refc["fstar"]) <= TOL and nrel(probs, refc["probs"]) <= TOL
    assert np.array_equal(np.asarray(gpc.df["predicted_class"][te], dtype=np.float64), refc["labels"])
    out = capsys.readouterr().out
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        assert out.count("matplotlib not available: plot skipped") >= 9


@pytest.mark.parametrize("kt,theta", [("diffusion", np.log([0.6, 0.1])), ("regularized_laplacian", np.log([0.6, 0.1])),
                                      ("pstep_walk", np.log([2.1, 2, 0.1])), ("pstep_walk", np.log([2.5, 1.5]))])
def test_classifier_gradient_matches_oracle(kt, theta):
    """GPnetClassifier.gradLogPosterior (RW Alg. 5.1 as written at GPnetClassifier.py:149-182; IndexError in the reference
    itself, A-12): the device path against the oracle's line-by-line restatement fed by the same derivative kernels."""
    import gpnet_b200
    G, tr, te, y = cfg.ba_case(300, 150, 100)
    ys = pd.Series(y[tr], index=tr)
    m = gpnet_b200.GPnetClassifier(Graph=G, training_nodes=tr, test_nodes=te, theta=list(theta), relabel_nodes=True, kerneltype=kt,
                                   training_values=ys)
    g = m.gradLogPosterior(theta, tr, ys)
    val, ref = go.OracleGP(G, kerneltype=kt, expm_mode="dense").classifier_grad(theta, tr, y[tr])
    assert g.shape == (len(theta),)
    assert nrel(g, ref) <= 1e-8
    assert abs(float(np.squeeze(m.logPosterior(theta, tr, ys))) - val) <= TOL * abs(val)


def test_large_n_diffusion_semigroup(gpu_ctx):
    """C5 family (2-D grid, diffusion) at N = 16384 (2.1 GB per matrix; the full N = 32768 case is tools/c5_check.py):
    exp(-b L~) exp(-b L~) = exp(-2b L~) on a column sample, product formed on the device through the C-ABI GEMM."""
    import torch
    a, b = 128, 128
    N = a * b
    indptr, indices, w = cfg.grid_csr_fast(a, b)
    gpu_ctx.set_graph(indptr, indices, w)
    cols = np.arange(0, N, N // 32, dtype=np.int32)[:32]
    assert gpu_ctx.kernel_build(0, np.log([0.3, 0.01]), False) == 0
    ptr, ld = gpu_ctx.kernel_ptr()
    K1c = torch.from_numpy(np.ascontiguousarray(gpu_ctx.kernel_block(range(N), cols, 0.0))).cuda()
    P = torch.zeros((N, 32), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    gpu_ctx.gemm(0, 0, N, 32, N, 1.0, ptr, ld, K1c.data_ptr(), 32, 0.0, P.data_ptr(), 32)
    gpu_ctx.sync()
    assert gpu_ctx.kernel_build(0, np.log([0.6, 0.01]), False) == 0
    K2c = gpu_ctx.kernel_block(range(N), cols, 0.0)
    assert np.max(np.abs(P.cpu().numpy() - K2c)) < 1e-12


def test_notebook_known_answer_on_gpu(gpu_ctx):
    """The reference's only KAT (report_notebook.ipynb cell 49) through the CUDA Newton loop."""
    import torch
    g = load_golden(os.path.join(os.path.dirname(__file__), "golden", "notebook_kat.npz"))
    n = len(g["y"])
    Kb = torch.zeros(n, 48, dtype=torch.float64, device="cuda:0")
    Kb[:, :n] = torch.tensor(g["K"], device="cuda:0")
    torch.cuda.synchronize()
    st, r = gpu_ctx.gpc_newton(0, [0.0], g["y"], variant=1, K_ptr=Kb.data_ptr(), ldk=48)
    assert st == 0 and r["info"] == 0
    assert abs(-r["logq"] - (-77.22939013)) < 5e-9
    assert r["iters"] == 5
    assert nrel(r["f"], g["f"]) <= TOL and nrel(r["a"], g["a"]) <= TOL


def test_error_behaviour_matches_reference():
    import gpnet_b200
    G, tr, te, t = cfg.grid_case(8, 8, 20, 20)
    m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=tr, test_nodes=te, theta=list(np.log([1.5, 0.01])), relabel_nodes=True,
                                  kerneltype="diffusion", training_values=False)
    m.set_training_values(pd.Series(t[tr], index=tr))
    with pytest.raises(AssertionError, match="Lambda must be < 1"):        # GPnet.py:339
        m.kernel(tr, tr, m.theta)
    with pytest.raises(AssertionError, match="Lambda must be < 1"):
        m.logp()
    m.kerneltype = "pstep_walk"
    with pytest.raises(AssertionError, match="a must be >=2"):             # GPnet.py:344
        m.kernel(tr, tr, np.log([1.5, 2.0, 0.1]))


def test_is_pos_def_and_calc_ktot():
    import gpnet_b200
    G, tr, te, t = cfg.grid_case(6, 6, 10, 10)
    m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=tr, test_nodes=te, theta=list(np.log([0.5, 0.1])), relabel_nodes=True,
                                  kerneltype="regularized_laplacian")
    m.calc_ktot()
    orc = go.OracleGP(G, kerneltype="regularized_laplacian")
    assert nrel(m.ktot, orc.kernel(range(36), range(36), m.theta)) <= TOL
    assert m.is_pos_def(m.ktot)
    assert not m.is_pos_def(-np.eye(5))


def test_optimize_params_with_device_gradient():
    """scipy.optimize drives the fused logp+grad entry point (jac=True extension) and the reference-style
    function-only mode to the same optimum (reduced C3: RGG, regularized Laplacian, L-BFGS-B box)."""
    import gpnet_b200
    G, tr, te, t = cfg.rgg_case(600, 300, 300)
    bounds = [[np.log(0.01), np.log(0.99)], [np.log(0.001), np.log(10.0)]]
    res = {}
    for jac in (True, False):
        m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=tr, test_nodes=te, theta=list(np.log([0.6, 0.01])), relabel_nodes=True,
                                      kerneltype="regularized_laplacian", training_values=False,
                                      optimize={"method": "L-BFGS-B", "bounds": bounds, "jac": jac})
        m.set_training_values(pd.Series(t[tr], index=tr))
        m.optimize_params()
        res[jac] = (np.array(m.theta), m.optimize_result.fun)
    orc = go.OracleGP(G, kerneltype="regularized_laplacian")
    assert abs(res[True][1] - orc.neg_logp(res[True][0], tr, t[tr])) <= 1e-9 * abs(res[True][1])
    assert abs(res[True][1] - res[False][1]) <= 1e-5 * abs(res[False][1])


def test_full_size_properties_c3_reglap(gpu_ctx):
    """BASELINE size (N = 8192, regularized Laplacian): size-independent identities instead of an O(N^3) CPU oracle:
    K (I + beta L~) = I on random columns, symmetry, and logp+grad consistent with a central difference of logp."""
    import torch
    indptr, indices, w, pos, t = cfg.rgg_csr_fast(8192)
    gpu_ctx.set_graph(indptr, indices, w)
    tr, te = cfg.split(8192, 4096, 4096)
    gpu_ctx.set_nodes(tr, te)
    beta = 0.6
    th = np.log([beta, 0.01])
    assert gpu_ctx.kernel_build(1, th, False) == 0
    cols = np.sort(np.random.RandomState(0).choice(8192, 64, replace=False)).astype(np.int32)
    Kc = gpu_ctx.kernel_block(range(8192), cols, 0.0)                    # 8192 x 64 columns of K
    Lnorm = go.normalized_laplacian(indptr, indices, w, 8192)
    M = np.eye(8192) + beta * Lnorm
    R = M @ Kc
    E = np.zeros_like(R)
    E[cols, np.arange(64)] = 1.0
    assert np.max(np.abs(R - E)) < 1e-12
    Kr = gpu_ctx.kernel_block(cols, range(8192), 0.0)
    assert np.array_equal(Kr, Kc.T)
    tt = t[tr]
    st, out, info = gpu_ctx.gpr_logp_grad(1, th, tt, True)
    assert st == 0 and info == 0
    h = 1e-5
    for j in range(2):
        tp, tm = th.copy(), th.copy()
        tp[j] += h
        tm[j] -= h
        fd = (gpu_ctx.gpr_logp_grad(1, tp, tt, False)[1][0] - gpu_ctx.gpr_logp_grad(1, tm, tt, False)[1][0]) / (2 * h)
        assert abs(fd - out[1 + j]) <= 1e-6 * max(1.0, abs(fd))


@pytest.mark.parametrize("shape", [(20, 13, 100), (37, 29, 300), (40, 40, 800), (32, 32, 1023), (48, 43, 1100), (64, 48, 1536), (64, 64, 1700), (71, 67, 2100)])
def test_reglap_fast_path_equals_full_path(gpu_ctx, shape):
    """The training-nodes-last routes (1: only K[tr,:] formed; 2: Schur complement read off the Cholesky factor, no kernel
    block formed; 3: other nodes eliminated through the sparse coupling block; 4: the same with the triangular work appended
    to the two Cholesky kernels, used when both blocks have at least 256 rows) against the full-kernel route on the same inputs, ragged block boundaries included (N - n not a multiple
    of 128), with and without the gradient."""
    a, b, n = shape
    N = a * b
    indptr, indices, w = cfg.grid_csr_fast(a, b)
    gpu_ctx.set_graph(indptr, indices, w)
    tr, te = cfg.split(N, n, min(N - n, 50), seed=3)
    gpu_ctx.set_nodes(tr, te)
    t = np.sin(0.3 * np.arange(n))
    th = np.log([0.7, 0.05])
    res = {}
    try:
        gpu_ctx.set_dissection_parts(2 + (N % 3))        # level 6 on small graphs: 2, 3 or 4 parts
        for level in (6, 5, 4, 3, 2, 1, 0):
            gpu_ctx.set_fast_paths(level)
            res[level] = (gpu_ctx.gpr_logp_grad(1, th, t, True), gpu_ctx.gpr_logp_grad(1, th, t, False))
            if level == 6 and N >= 600:
                assert gpu_ctx.last_route()[0] == 6, (N, gpu_ctx.last_route())
    finally:
        gpu_ctx.set_fast_paths(True)
        gpu_ctx.set_dissection_parts(0)
        gpu_ctx.set_dissection_parts(0)
    (s0, g0, i0), (_, l0, _) = res[0]
    assert s0 == 0 and i0 == 0
    for level in (1, 2, 3, 4, 5, 6):
        (s1, g1, i1), (_, l1, _) = res[level]
        assert s1 == 0 and i1 == 0
        assert nrel(g1, g0) <= 1e-11, (level, g1, g0)
        assert abs(l1[0] - l0[0]) <= 1e-12 * abs(l0[0]) and abs(g1[0] - l1[0]) <= 1e-12 * abs(l1[0])


@pytest.mark.parametrize("beta,noise", [(0.01, 1e-6), (0.01, 100.0), (30.0, 1e-6), (30.0, 100.0), (0.99, 10.0)])
def test_reglap_routes_agree_with_oracle_for_stiff_parameters(gpu_ctx, beta, noise):
    """Every route of the regularized-Laplacian likelihood (full kernel ... Schur complement through the sparse coupling block
    with augmented factorisations) against the numpy oracle at the ends of the parameter box and beyond (cond(I + beta L~) up
    to 61, noise from 1e-6 to 100): the rank-one / Schur algebra must not lose digits where K_y is dominated by either term."""
    G, tr, te, t = cfg.rgg_case(700, 300, 100)
    indptr, indices, w, _ = go.graph_to_csr(G)
    gpu_ctx.set_graph(indptr, indices, w)
    gpu_ctx.set_nodes(tr, te)
    th = np.log([beta, noise])
    orc = go.OracleGP(G, kerneltype="regularized_laplacian")
    ref_v, ref_g = orc.neg_logp_grad(th, tr, t[tr])
    try:
        gpu_ctx.set_dissection_parts(3)
        for level in (0, 1, 2, 3, 4, 5, 6):
            gpu_ctx.set_fast_paths(level)
            st, out, info = gpu_ctx.gpr_logp_grad(1, th, t[tr], True)
            assert st == 0 and info == 0
            assert abs(out[0] - ref_v) <= TOL * abs(ref_v), (level, out[0], ref_v)
            assert nrel(out[1:], ref_g) <= 1e-8, (level, out[1:], ref_g)
    finally:
        gpu_ctx.set_fast_paths(True)
        gpu_ctx.set_dissection_parts(0)


@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15, 16])
def test_reglap_routes_random_shapes_against_oracle(gpu_ctx, seed):
    """Random geometric graphs of random size, training fraction, beta and noise (ragged blocks on both sides of the 256-row
    threshold of the augmented factorisations): every route against the numpy oracle (tools/fuzz_routes.py is the long form)."""
    rng = np.random.RandomState(seed)
    N = int(rng.randint(300, 2200))
    n = max(2, int(N * rng.uniform(0.12, 0.88)))
    beta = float(np.exp(rng.uniform(np.log(0.01), np.log(5.0))))
    noise = float(np.exp(rng.uniform(np.log(1e-4), np.log(10.0))))
    G = nx.random_geometric_graph(N, np.sqrt(rng.uniform(6, 24) / (np.pi * N)), seed=int(rng.randint(1 << 30)))
    perm = rng.permutation(N)
    tr = sorted(int(x) for x in perm[:n])
    te = sorted(int(x) for x in perm[n:n + min(N - n, 16)])
    t = rng.randn(n)
    indptr, indices, w, _ = go.graph_to_csr(G)
    gpu_ctx.set_graph(indptr, indices, w)
    gpu_ctx.set_nodes(tr, te)
    th = np.log([beta, noise])
    ref_v, ref_g = go.OracleGP(G, kerneltype="regularized_laplacian").neg_logp_grad(th, tr, t)
    try:
        gpu_ctx.set_dissection_parts(3)
        for level in (0, 1, 2, 3, 4, 5, 6):
            gpu_ctx.set_fast_paths(level)
            st, out, info = gpu_ctx.gpr_logp_grad(1, th, t, True)
            assert st == 0 and info == 0
            assert abs(out[0] - ref_v) <= TOL * abs(ref_v), (level, N, n, out[0], ref_v)
            assert nrel(out[1:], ref_g) <= 1e-8, (level, N, n, out[1:], ref_g)
    finally:
        gpu_ctx.set_fast_paths(True)
        gpu_ctx.set_dissection_parts(0)


def test_reglap_routes_on_a_weighted_graph_with_isolated_nodes_and_a_training_only_component(gpu_ctx):
    """Edge cases of the Schur routes' graph handling: random edge weights (the sparse coupling block carries them), isolated
    nodes on both sides (zero degree: their row of I + beta L~ is the identity), a component made of training nodes only (no
    coupling at all) and training nodes without any other-node neighbour -- all five routes against the numpy oracle."""
    rng = np.random.RandomState(7)
    G = nx.random_geometric_graph(640, 0.09, seed=3)
    G.add_nodes_from(range(640, 700))                              # 60 isolated nodes
    for a in range(660, 680):                                      # a path component among some of them
        G.add_edge(a, a + 1)
    for (u, v) in G.edges():
        G[u][v]["weight"] = float(rng.uniform(0.2, 3.0))
    N = G.number_of_nodes()
    perm = rng.permutation(640)
    tr = sorted([int(x) for x in perm[:300]] + list(range(650, 685)))      # the path component and some isolated nodes are training nodes
    te = sorted(int(x) for x in perm[300:360])
    t = np.sin(0.05 * np.arange(len(tr))) + 0.1 * rng.randn(len(tr))
    indptr, indices, w, _ = go.graph_to_csr(G)
    gpu_ctx.set_graph(indptr, indices, w)
    gpu_ctx.set_nodes(tr, te)
    th = np.log([0.8, 0.05])
    orc = go.OracleGP(G, kerneltype="regularized_laplacian")
    ref_v, ref_g = orc.neg_logp_grad(th, tr, t)
    try:
        gpu_ctx.set_dissection_parts(3)
        for level in (0, 1, 2, 3, 4, 5, 6):
            gpu_ctx.set_fast_paths(level)
            st, out, info = gpu_ctx.gpr_logp_grad(1, th, t, True)
            assert st == 0 and info == 0
            assert abs(out[0] - ref_v) <= TOL * abs(ref_v), (level, out[0], ref_v)
            assert nrel(out[1:], ref_g) <= 1e-8, (level, out[1:], ref_g)
    finally:
        gpu_ctx.set_fast_paths(True)
        gpu_ctx.set_dissection_parts(0)


@pytest.mark.parametrize("kind,theta,axidx", [(0, np.log([0.6, 0.01]), [0, 1]), (0, np.log([0.6, 0.01]), [1, 0]),
                                              (1, np.log([0.6, 0.01]), [0, 1]), (2, np.log([2.3, 4, 0.01]), [0, 2]),
                                              (2, np.log([2.3, 4]), [0, 1]), (1, np.log([0.6, 0.01]), [1, 1])])
@pytest.mark.parametrize("classifier", [False, True])
def test_landscape_kernel_reuse_equals_independent_points(gpu_ctx, kind, theta, axidx, classifier):
    """gpn_landscape_points builds the kernel once per group of points that differ only in the noise entry (regressor: one
    factorisation of K_tt, every noise value as a rank-one update): same grid as one independent evaluation per point
    (fast paths off), in any index order, with the gradient."""
    indptr, indices, w = cfg.grid_csr_fast(17, 19)
    N = 17 * 19
    gpu_ctx.set_graph(indptr, indices, w)
    tr, te = cfg.split(N, 140, 60, seed=5)
    gpu_ctx.set_nodes(tr, te)
    rng = np.random.RandomState(2)
    t = np.sign(rng.randn(140)) if classifier else np.sin(0.2 * np.arange(140))
    lo0 = {0: np.log(0.05), 1: np.log(0.05), 2: np.log(2.0)}[kind]
    hi0 = {0: np.log(0.95), 1: np.log(3.0), 2: np.log(3.0)}[kind]
    axes = {0: np.linspace(lo0, hi0, 4), 1: np.linspace(np.log(1e-3), np.log(5.0), 5), 2: np.linspace(np.log(1e-3), np.log(5.0), 5)}
    if kind == 2 and len(theta) == 2:
        axes[1] = np.log(np.array([1.2, 2.0, 3.5, 4.1, 6.0]))        # the order p AND the noise (A-4): no sharing possible
    a1, a2 = axes[axidx[0]], axes[axidx[1]]
    idx = rng.permutation(len(a1) * len(a2))
    res = {}
    try:
        for level in (5, 0):
            gpu_ctx.set_fast_paths(level)
            st, out = gpu_ctx.landscape_points(kind, classifier, theta, axidx, a1, a2, t, idx, want_grad=not classifier)
            assert st == 0
            res[level] = out
    finally:
        gpu_ctx.set_fast_paths(True)
        gpu_ctx.set_dissection_parts(0)
    assert np.all(np.isfinite(res[0]))
    assert nrel(res[5][:, 0], res[0][:, 0]) <= 1e-11
    if not classifier:
        assert nrel(res[5][:, 1:], res[0][:, 1:]) <= 1e-10
    # the strided form of the same call covers the same points
    st, out = gpu_ctx.landscape(kind, classifier, theta, axidx, a1, a2, t, 1, len(idx), 3, want_grad=False)
    ref = {int(i): v for i, v in zip(idx, res[0][:, 0])}
    assert st == 0 and nrel(out[:, 0], np.array([ref[i] for i in range(1, len(idx), 3)])) <= 1e-11


def test_diffusion_semigroup_property_n4096(gpu_ctx):
    """C4 size: exp(-b L~) exp(-b L~) = exp(-2b L~) on a column sample (Pade orders differ between the two sides)."""
    indptr, indices, w, pos, t = cfg.rgg_csr_fast(4096)
    gpu_ctx.set_graph(indptr, indices, w)
    cols = np.arange(0, 4096, 128, dtype=np.int32)
    assert gpu_ctx.kernel_build(0, np.log([0.45, 0.01]), False) == 0
    m1 = gpu_ctx.expm_info()
    K1 = gpu_ctx.kernel_block(range(4096), range(4096), 0.0)
    assert gpu_ctx.kernel_build(0, np.log([0.9, 0.01]), False) == 0
    m2 = gpu_ctx.expm_info()
    K2c = gpu_ctx.kernel_block(range(4096), cols, 0.0)
    assert np.max(np.abs(K1 @ K1[:, cols] - K2c)) < 1e-12
    assert m1[0] in (5, 7, 9, 13) and m2[0] in (7, 9, 13)


# ---------------------------------------------------------------------------------------------------------------------
# Full-size parity against the oracle: the headline configuration itself, the default route AND the
# reference's order of operations, with the route that actually ran asserted.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c3_case():
    """C3 of SURVEY Appendix D at full size: nx.random_geometric_graph(8192, sqrt(20/(pi N)), seed=0), 4096/4096 split,
    t = sin(4 pi x) cos(4 pi y), theta = log[0.6, 0.01]; oracle value + analytic gradient computed once (dense, ~10 s)."""
    G, tr, te, t = cfg.rgg_case(8192, 4096, 4096)
    assert nx.is_connected(G)
    th = np.log([0.6, 0.01])
    ref_v, ref_g = go.OracleGP(G, kerneltype="regularized_laplacian").neg_logp_grad(th, tr, t[tr])
    return G, tr, te, t, th, ref_v, ref_g


@pytest.mark.parametrize("level", [None, 6, 5, 4, 2, 0])
def test_c3_full_size_logp_and_gradient_against_oracle(c3_case, level):
    """GPnetRegressor.logPosterior_and_grad at N = 8192 / n = 4096 (GPnetRegressor.py:136-150 + the gradient of
    old_implementation/GPR.py:52-64) against OracleGP.neg_logp_grad: value to 1e-9, gradient to 1e-8, on the default route
    (level None), the dissected / augmented Schur routes, and level 0 = the reference's own order of operations.  The route
    that ran is read back (gpn_last_route) so that a silent fallback fails the test."""
    import gpnet_b200
    G, tr, te, t, th, ref_v, ref_g = c3_case
    m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=tr, test_nodes=te, theta=list(th), relabel_nodes=True,
                                  kerneltype="regularized_laplacian", training_values=False)
    tv = pd.Series(t[tr], index=tr)
    m.set_training_values(tv)
    ctx = m._engine.ctx
    try:
        if level is not None:
            ctx.set_fast_paths(level)
        val, grad = m.logPosterior_and_grad(th, tr, tv)
        route, parts = ctx.last_route()
        val_only = m.logPosterior(th, tr, tv)
        route_v, _ = ctx.last_route()
    finally:
        ctx.set_fast_paths(True)
    assert abs(val - ref_v) <= TOL * abs(ref_v), (level, val, ref_v)
    assert nrel(grad, ref_g) <= 1e-8, (level, grad, ref_g)
    assert abs(val_only - ref_v) <= TOL * abs(ref_v)
    if level is None:
        assert route >= 5 and parts[0] >= 2 and parts[1] > 0, (route, parts)      # the dissection really ran on the headline graph
        assert route_v >= 5
    else:
        assert route == level, (route, level)
        if level == 5:
            assert parts[0] == 2 and min(parts[2:4]) >= 512 and 0 < parts[1] <= min(parts[2:4]) // 2
        if level == 6:
            K = parts[0]
            assert K >= 4 and min(parts[2:2 + K]) >= 256 and sum(parts[1:2 + K]) == 8192 and parts[1] < 1024


@pytest.mark.parametrize("N,n,seed", [(4096, 1500, 1), (5000, 2500, 2), (6000, 2000, 3)])
def test_dissected_route_against_oracle_on_graphs_with_many_other_nodes(gpu_ctx, N, n, seed):
    """Random geometric graphs whose other-node block is large enough (o >= 2048) for the level-set dissection: every route
    against the dense oracle, and the dissection must really have been taken on the default level."""
    rng = np.random.RandomState(seed)
    G = nx.random_geometric_graph(N, np.sqrt(18.0 / (np.pi * N)), seed=seed)
    perm = rng.permutation(N)
    tr = sorted(int(x) for x in perm[:n])
    te = sorted(int(x) for x in perm[n:n + 32])
    t = rng.randn(n)
    indptr, indices, w, _ = go.graph_to_csr(G)
    gpu_ctx.set_graph(indptr, indices, w)
    gpu_ctx.set_nodes(tr, te)
    th = np.log([0.8, 0.02])
    ref_v, ref_g = go.OracleGP(G, kerneltype="regularized_laplacian").neg_logp_grad(th, tr, t)
    try:
        for level in (True, 6, 5, 4, 0):
            gpu_ctx.set_fast_paths(level)
            st, out, info = gpu_ctx.gpr_logp_grad(1, th, t, True)
            route, parts = gpu_ctx.last_route()
            st2, out2, info2 = gpu_ctx.gpr_logp_grad(1, th, t, False)
            assert st2 == 0 and info2 == 0 and abs(out2[0] - ref_v) <= TOL * abs(ref_v), (level, out2[0], ref_v)
            assert st == 0 and info == 0
            assert abs(out[0] - ref_v) <= TOL * abs(ref_v), (level, out[0], ref_v)
            assert nrel(out[1:], ref_g) <= 1e-8, (level, out[1:], ref_g)
            if level is True or level >= 5:
                assert route == (6 if level is True else level) and parts[0] >= 2 and parts[1] > 0, (level, route, parts)
            else:
                assert route == level
    finally:
        gpu_ctx.set_fast_paths(True)
        gpu_ctx.set_dissection_parts(0)


def test_c4_landscape_points_against_dense_oracle():
    """C4 (RGG N = 4096, n = 2048, diffusion): three points of the 64 x 64 landscape (Pade orders 5 / 7 / 9) through
    GPnetRegressor.lml_landscape sub-grids against the dense oracle (GPnet.py:539-552)."""
    import gpnet_b200
    G, tr, te, t = cfg.rgg_case(4096, 2048, 0)
    ax1 = np.linspace(np.log(0.01), np.log(0.99), 64)
    ax2 = np.linspace(np.log(0.001), np.log(10.0), 64)
    m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=tr, test_nodes=[], theta=list(np.log([0.6, 0.01])), relabel_nodes=True,
                                  kerneltype="diffusion", training_values=False)
    tv = pd.Series(t[tr], index=tr)
    m.set_training_values(tv)
    orc = go.OracleGP(G, kerneltype="diffusion", expm_mode="dense")
    orders = set()
    for i1, i2 in ((20, 5), (45, 30), (63, 63)):
        th = [0.0, 0.0]
        lml = m.lml_landscape(th, [0, 1], ax1[i1:i1 + 1], ax2[i2:i2 + 1])
        orders.add(m._engine.ctx.expm_info()[0])
        ref = -orc.neg_logp([ax1[i1], ax2[i2]], tr, t[tr])
        assert lml.shape == (1, 1) and abs(lml[0, 0] - ref) <= TOL * abs(ref), (i1, i2, lml, ref)
        val, grad = m.logPosterior_and_grad([ax1[i1], ax2[i2]], tr, tv)
        assert abs(-val - ref) <= TOL * abs(ref)
    assert len(orders) >= 2          # the three points exercise different Pade orders


def test_c5_family_grid_4096_against_dense_oracle():
    """C5 family (2-D grid, diffusion, t = sin(0.05 d0)) at N = 4096 = 64 x 64, n = m = 2048: predict(), logp and the
    analytic gradient against the dense oracle (the full N = 32768 case is tools/c5_check.py: identities only)."""
    import gpnet_b200
    G, tr, te, t = cfg.grid_case(64, 64, 2048, 2048, freq=0.05)
    th = np.log([0.6, 0.01])
    m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=tr, test_nodes=te, theta=list(th), relabel_nodes=True,
                                  kerneltype="diffusion", training_values=False)
    tv = pd.Series(t[tr], index=tr)
    m.set_training_values(tv)
    orc = go.OracleGP(G, kerneltype="diffusion", expm_mode="dense")
    pr = orc.predict_regressor(th, tr, te, t[tr])
    assert m.predict() is m and m.is_trained
    assert nrel(m.fstar, pr["fstar"]) <= TOL and nrel(m.s, pr["s"]) <= TOL and nrel(m.alpha, pr["alpha"]) <= TOL
    ref_v, ref_g = orc.neg_logp_grad(th, tr, t[tr])
    val, grad = m.logPosterior_and_grad(th, tr, tv)
    assert abs(val - ref_v) <= TOL * abs(ref_v) and nrel(grad, ref_g) <= 1e-8
    # the lazily fetched attributes survive the logPosterior call above (device buffers were reused)
    assert nrel(m.kstar, pr["kstar"]) <= TOL and nrel(m.k, pr["k"]) <= TOL
    L = m.L
    assert nrel(L @ L.T, pr["k"]) <= TOL
    assert m.v.shape == (2048, 2048)


def test_kernel_with_labels_outside_the_graph_raises_index_error():
    """GPnetBase.kernel with a label that is no graph position must not read out of bounds on the device."""
    import gpnet_b200
    G, tr, te, t = cfg.grid_case(8, 8, 20, 20)
    m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=tr, test_nodes=te, theta=list(np.log([0.5, 0.01])), relabel_nodes=True,
                                  kerneltype="diffusion")
    with pytest.raises(IndexError):
        m.kernel([0, 1, 64], [0], m.theta)
    with pytest.raises(IndexError):
        m.kernel([0], [-3], m.theta)
    assert m.kernel([0, 1], [2], m.theta).shape == (2, 1)
