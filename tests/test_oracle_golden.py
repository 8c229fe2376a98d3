"""CPU tier: the numpy oracle against the golden vectors produced by the UNMODIFIED (shimmed) reference
(oracle/make_golden.py), the reference's only known-answer value, and independent cross-checks."""
import os

import numpy as np
import pytest

from conftest import golden_files, load_golden, nrel
from oracle import gp_oracle as go

TOL = 1e-9   # north_star tolerance (norm-wise relative); the oracle itself agrees to ~1e-14


def _orc(g, expm_mode="sparse"):
    return go.OracleGP(csr=(g["indptr"], g["indices"], g["weights"]), kerneltype=str(g["kerneltype"]), expm_mode=expm_mode)


@pytest.mark.parametrize("path", golden_files("reg_"), ids=os.path.basename)
def test_regressor_oracle_matches_reference(path):
    g = load_golden(path)
    N = len(g["indptr"]) - 1
    orc = _orc(g, "sparse" if N <= 1100 else "dense")
    tr, te, t, th = g["train"], g["test"], g["t_train"], g["theta"]
    assert abs(-orc.neg_logp(th, tr, t) - g["logp"]) <= TOL * abs(g["logp"])
    pr = orc.predict_regressor(th, tr, te, t)
    for key in ("fstar", "s", "alpha", "kstarstar_diag"):
        assert nrel(pr[key], g[key]) <= TOL, key
    if "k" in g:
        assert nrel(pr["k"], g["k"]) <= TOL and nrel(pr["kstar"], g["kstar"]) <= TOL
    if "neg_logp_grad_fd" in g:     # analytic gradient vs central differences of the reference's logPosterior
        _, ng = orc.neg_logp_grad(th, tr, t)
        sel = g["grad_fd_valid"]
        assert nrel(ng[sel], g["neg_logp_grad_fd"][sel]) <= 1e-6
    if "land_lml" in g:
        lml = orc.landscape(th, list(g["land_axidx"]), g["land_ax1"], g["land_ax2"], tr, t)
        assert nrel(lml, g["land_lml"]) <= TOL


@pytest.mark.parametrize("path", golden_files("cls_"), ids=os.path.basename)
def test_classifier_oracle_matches_reference(path):
    g = load_golden(path)
    N = len(g["indptr"]) - 1
    orc = _orc(g, "sparse" if N <= 500 else "dense")
    pr = orc.predict_classifier(g["theta"], g["train"], g["test"], g["y_train"])
    assert pr["iters"] == g["iters"]
    assert abs(pr["logq"] - g["logq"]) <= TOL * abs(g["logq"])
    for key in ("f", "a", "fstar", "V", "probs"):
        assert nrel(np.asarray(pr[key]).reshape(np.asarray(g[key]).shape), g[key]) <= TOL, key
    assert np.array_equal(pr["labels"], g["labels"])


def test_notebook_known_answer():
    """report_notebook.ipynb cell 49: logPosterior = -77.22939013 (old-variant logq), 5 Newton iterations."""
    K, y = go.notebook_kat_inputs()
    f, logq, a, iters = go.gpc_newton(K, y, logq_variant="notebook")
    assert abs(-logq - go.NOTEBOOK_KAT_VALUE) < 5e-9
    assert iters == 5
    g = load_golden(os.path.join(os.path.dirname(__file__), "golden", "notebook_kat.npz"))
    assert np.allclose(K, g["K"]) and np.array_equal(y.ravel(), g["y"])


@pytest.mark.parametrize("kt,theta", [("diffusion", np.log([0.6, 0.01])), ("regularized_laplacian", np.log([0.6, 0.01]))])
def test_spectral_cross_check(kt, theta):
    """K = V r(lambda) V^T (PDF eq. 55-56) agrees with the expm / inverse paths."""
    import networkx as nx
    G = nx.relabel_nodes(nx.grid_2d_graph(12, 11), dict(zip(nx.grid_2d_graph(12, 11), range(132))))
    a = go.OracleGP(G, kerneltype=kt, expm_mode="sparse").full(theta)
    b = go.OracleGP(G, kerneltype=kt, expm_mode="spectral").full(theta)
    c = go.OracleGP(G, kerneltype=kt, expm_mode="dense").full(theta)
    assert nrel(a, b) < 1e-12 and nrel(c, b) < 1e-12


def test_kernel_properties_and_quirks():
    import networkx as nx
    G = nx.barabasi_albert_graph(60, 3, seed=1)
    orc = go.OracleGP(G, kerneltype="regularized_laplacian")
    K = orc.full(np.log([0.7, 0.1]))
    assert np.allclose(K @ (np.eye(60) + 0.7 * orc.Lnorm), np.eye(60), atol=1e-12)
    orc.kerneltype = "pstep_walk"
    assert np.allclose(orc.full(np.log([2.5, 1.0, 0.1])), 2.5 * np.eye(60) - orc.Lnorm)
    # A-4: p = int(exp(log 5)) == 4
    assert go.pstep_order(np.exp(np.log([2.5, 5.0]))) == 4
    # A-3: noise on every entry; A-5: graph-position order regardless of caller order
    blk = orc.kernel([5, 2, 9], [7, 1], np.log([2.5, 2.0, 0.3]))
    Kf = orc.full(np.log([2.5, 2.0, 0.3]))
    assert np.allclose(blk, Kf[np.ix_([2, 5, 9], [1, 7])] + 0.3)
    # parameter-domain asserts (GPnet.py:339,344)
    orc.kerneltype = "diffusion"
    with pytest.raises(AssertionError):
        orc.full(np.log([1.5, 0.1]))
    orc.kerneltype = "pstep_walk"
    with pytest.raises(AssertionError):
        orc.full(np.log([1.5, 2.0, 0.1]))


def test_laplacian_matches_networkx_edge_cases():
    import networkx as nx
    G = nx.path_graph(7)
    G.add_edge(2, 2)
    G.add_node(7)
    G[0][1]["weight"] = 2.5
    indptr, indices, w, _ = go.graph_to_csr(G)
    L = go.normalized_laplacian(indptr, indices, w, 8)
    assert np.array_equal(L, nx.normalized_laplacian_matrix(G).toarray())


def test_neg_inf_on_cholesky_failure():
    assert go.gpr_neg_logp(-np.eye(3), np.ones(3)) == -np.inf


@pytest.mark.parametrize("N,n,beta,c", [(420, 180, 0.6, 0.01), (333, 129, 0.9, 3.0), (300, 290, 0.05, 1e-3), (256, 1, 0.6, 0.01)])
def test_schur_route_algebra_matches_the_oracle(N, n, beta, c):
    """The algebra of the default regularized-Laplacian route of the CUDA path (csrc/gp.cu
    gpr_reglap_coupled_scalars + row_eval), restated in numpy and checked against the oracle's Cholesky-of-K_y evaluation:
    training nodes last, Z = M_oo^-1, X = M_to Z, S = M_tt - X M_ot = L L^T = K_tt^-1, the all-entries noise as a rank-one
    update, and the gradient traces from Y = X^T L^-T and W = L^-1 (||Y||_F^2 + ||W||_F^2 - n = tr(S dK_tt/dtheta_0))."""
    import scipy.linalg as sl
    from oracle import configs as cfg
    rng = np.random.RandomState(N + n)
    indptr, indices, w = cfg.rgg_csr_fast(N)[:3]
    Lt = go.normalized_laplacian(indptr, indices, w, N)
    theta = np.log([beta, c])
    tr = np.sort(rng.permutation(N)[:n])
    t = rng.randn(n)
    K = go.full_kernel(Lt, "regularized_laplacian", theta)
    Ky = go.kernel_block(K, tr, tr, theta, 1.0)
    dK0 = go.kernel_derivatives(Lt, K, "regularized_laplacian", theta)[np.ix_(tr, tr)]
    ref = np.concatenate([[go.gpr_neg_logp(Ky, t)], go.gpr_neg_logp_grad(Ky, [dK0, c * np.ones((n, n))], t)])

    oth = np.setdiff1d(np.arange(N), tr)
    o = N - n
    M = np.eye(N) + beta * Lt
    Moo, Mto, Mtt = M[np.ix_(oth, oth)], M[np.ix_(tr, oth)], M[np.ix_(tr, tr)]
    Z = np.linalg.inv(Moo)
    X = Mto @ Z                                   # sparse x dense on the device
    S = Mtt - X @ Mto.T                           # = K_tt^-1
    L = np.linalg.cholesky(S)
    r1, rt = L.T @ np.ones(n), L.T @ t
    s, a, q = r1 @ r1, r1 @ rt, rt @ rt
    den = 1.0 + c * s
    mu = c * a / den
    nlogp = 0.5 * (q - mu * a) + (-np.log(np.diag(L)).sum() + 0.5 * np.log1p(c * s)) + 0.5 * n * np.log(2 * np.pi)
    W = sl.solve_triangular(L, np.eye(n), lower=True)
    Y = sl.solve_triangular(L, X, lower=True).T   # X^T L^-T = -W_to^T
    z1 = np.concatenate([Y @ r1, np.ones(n)])     # K[:, tr] (S 1): the training part is exactly 1
    zt = np.concatenate([Y @ rt, t])              # K[:, tr] (S t): the training part is exactly t
    assert np.allclose(Y @ r1, X.T @ np.ones(n), rtol=1e-10, atol=1e-12) and np.allclose(Y @ rt, X.T @ t, rtol=1e-10, atol=1e-12)
    D11, D1t, Dtt = z1 @ z1 - s, z1 @ zt - a, zt @ zt - q
    trSD = (Y * Y).sum() + (W * W).sum() - n
    g0 = 0.5 * (Dtt - 2 * mu * D1t + mu * mu * D11) - 0.5 * (trSD - c / den * D11)
    gl = 0.5 * c * ((a / den) ** 2 - s / den)
    mine = np.array([nlogp, -g0, -gl])
    assert np.max(np.abs(mine - ref) / np.abs(ref)) < 1e-10


@pytest.mark.parametrize("kerneltype,theta0", [("diffusion", np.log(0.6)), ("pstep_walk", np.log(2.3))])
def test_landscape_row_algebra_matches_the_oracle(kerneltype, theta0):
    """What gpn_landscape_points does for grid points that differ only in the noise entry (csrc/gp.cu gpr_row_scalars +
    row_eval): ONE factorisation of K_tt, every noise value c as the rank-one update K_y = K_tt + c 1 1^T -- restated in numpy
    and checked against the oracle's independent evaluation of each point, value and gradient."""
    import scipy.linalg as sl
    from oracle import configs as cfg
    N, n = 260, 110
    rng = np.random.RandomState(3)
    indptr, indices, w = cfg.rgg_csr_fast(N)[:3]
    Lt = go.normalized_laplacian(indptr, indices, w, N)
    tr = np.sort(rng.permutation(N)[:n])
    t = rng.randn(n)
    P = 3 if kerneltype == "pstep_walk" else 2
    base = [theta0, np.log(4.0), 0.0] if P == 3 else [theta0, 0.0]
    K = go.full_kernel(Lt, kerneltype, np.array(base))
    Ktt = K[np.ix_(tr, tr)]
    D = go.kernel_derivatives(Lt, K, kerneltype, np.array(base))[np.ix_(tr, tr)]
    L = np.linalg.cholesky(Ktt)
    Wi = sl.solve_triangular(L, np.eye(n), lower=True)
    u, v = Wi @ np.ones(n), Wi @ t
    s, a, q = u @ u, u @ v, v @ v
    Sinv = Wi.T @ Wi
    x1, xt = Wi.T @ u, Wi.T @ v
    trSD, D11, D1t, Dtt = np.sum(Sinv * D), x1 @ D @ x1, x1 @ D @ xt, xt @ D @ xt
    half_logdet = np.log(np.diag(L)).sum()
    for c in (1e-4, 0.01, 0.7, 9.0):
        th = np.array(base)
        th[-1] = np.log(c)
        Ky = go.kernel_block(K, tr, tr, th, 1.0)
        dKs = [D] + [np.zeros((n, n))] * (P - 2) + [c * np.ones((n, n))]
        ref = np.concatenate([[go.gpr_neg_logp(Ky, t)], go.gpr_neg_logp_grad(Ky, dKs, t)])
        den = 1.0 + c * s
        mu = c * a / den
        val = 0.5 * (q - mu * a) + (half_logdet + 0.5 * np.log1p(c * s)) + 0.5 * n * np.log(2 * np.pi)
        g0 = 0.5 * (Dtt - 2 * mu * D1t + mu * mu * D11) - 0.5 * (trSD - c / den * D11)
        gl = 0.5 * c * ((a / den) ** 2 - s / den)
        mine = np.concatenate([[val, -g0], np.zeros(P - 2), [-gl]])
        assert np.max(np.abs(mine - ref) / np.maximum(np.abs(ref), 1e-300)) < 1e-9 or np.allclose(mine, ref, rtol=1e-9, atol=1e-12)
