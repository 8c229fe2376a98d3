"""numpy/scipy restatement of the reference's GP-on-graph hot path (CPU oracle; see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to the reference repo root,
filippocastelli/GPnet).  The arithmetic itself lives in un-vendored third-party dependencies of the
reference (``README.md:17`` names them without versions); the versions this oracle is pinned
against are numpy 2.3.5, scipy 1.18.1, networkx 3.6.1, pandas 3.0.2:

* ``networkx.normalized_laplacian_matrix`` (called at ``GPnet.py:329``) -> ``normalized_laplacian``;
* ``scipy.linalg.expm`` on a sparse matrix (``GPnet.py:340``; 2018 scipy delegated this to
  ``scipy.sparse.linalg.expm``, Al-Mohy & Higham 2009) -> ``full_kernel(..., expm_mode="sparse")``;
* ``scipy.linalg.inv`` (``GPnet.py:342``), ``numpy.linalg.matrix_power`` (``GPnet.py:346``),
  ``numpy.linalg.cholesky/solve/inv`` (``GPnetRegressor.py:123-126,141-144``,
  ``GPnetClassifier.py:110-126,208-211``), ``scipy.special.erf`` (``GPnetClassifier.py:228``).

Reference quirks reproduced on purpose: the noise ``exp(theta[-1])`` is added
to EVERY entry of the block (A-3); ``p = int(exp(theta[1]))`` (A-4); blocks come out in graph-position
order (A-5); ``logPosterior`` uses un-shifted targets while ``predict`` uses mean-shifted ones (A-6);
``-inf`` on Cholesky failure (A-8); the log-of-log ``logq`` and the loose n-vector convergence test of
the Newton loop (A-10, A-11).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg
from scipy.special import erf

# GPnet.py:59-62 (constants of the 5-erf approximation of the logistic sigmoid)
LAMBDAS = np.array([0.41, 0.4, 0.37, 0.44, 0.39])
COEFS = np.array([-1854.8214151, 3516.89893646, 221.29346712, 128.12323805, -2010.49422654])[:, np.newaxis]

KERNEL_TYPES = ("diffusion", "regularized_laplacian", "pstep_walk")


# ----------------------------------------------------------------------------------------------
# graph -> CSR -> normalized Laplacian            (GPnet.py:329, networkx laplacianmatrix.py body)
# ----------------------------------------------------------------------------------------------
def graph_to_csr(G, weight="weight"):
    """Adjacency of a networkx graph in CSR form with node order ``list(G)`` (GPnet.py:329 passes no
    nodelist, so networkx uses ``list(G)``).  Undirected edges appear in both rows; a self-loop
    appears once on the diagonal (networkx ``to_scipy_sparse_array`` convention)."""
    nodelist = list(G)
    index = {u: i for i, u in enumerate(nodelist)}
    N = len(nodelist)
    rows = [[] for _ in range(N)]
    for u, v, d in G.edges(data=True):
        w = float(d.get(weight, 1.0))
        i, j = index[u], index[v]
        rows[i].append((j, w))
        if i != j:
            rows[j].append((i, w))
    indptr = np.zeros(N + 1, dtype=np.int32)
    indices, weights = [], []
    for i, r in enumerate(rows):
        r.sort()
        # parallel edges cannot occur in nx.Graph; keep duplicates summed for safety
        merged = {}
        for j, w in r:
            merged[j] = merged.get(j, 0.0) + w
        for j in sorted(merged):
            indices.append(j)
            weights.append(merged[j])
        indptr[i + 1] = len(indices)
    return indptr, np.asarray(indices, dtype=np.int32), np.asarray(weights, dtype=np.float64), nodelist


def normalized_laplacian(indptr, indices, weights, N):
    """Dense ``L~ = D^-1/2 (D - A) D^-1/2`` with networkx's operation order
    (``DH @ ((D - A) @ DH)``; ``1/sqrt(d)`` with inf -> 0), GPnet.py:329."""
    A = scipy.sparse.csr_array((weights, indices, indptr), shape=(N, N))
    d = np.asarray(A.sum(axis=1)).ravel()
    with np.errstate(divide="ignore"):
        dinv = 1.0 / np.sqrt(d)
    dinv[np.isinf(dinv)] = 0.0
    L = (scipy.sparse.diags_array(d) - A).toarray()
    return dinv[:, None] * (L * dinv[None, :])


# ----------------------------------------------------------------------------------------------
# full N x N Kondor kernels                                                     (GPnet.py:335-349)
# ----------------------------------------------------------------------------------------------
def pstep_order(theta_exp):
    """``int(theta[1])`` after ``theta = exp(theta)`` (GPnet.py:303,347; quirk A-4)."""
    return int(theta_exp[1])


def full_kernel(Lnorm, kerneltype, theta, expm_mode="sparse"):
    """The N x N kernel before block selection.  ``theta`` is the LOG-parameter vector
    (``theta = np.exp(theta)`` at GPnet.py:303).  Raises AssertionError exactly where the reference
    does (GPnet.py:337,339,344)."""
    th = np.exp(np.squeeze(np.asarray(theta, dtype=np.float64)))
    th = np.atleast_1d(th)
    N = Lnorm.shape[0]
    assert kerneltype in KERNEL_TYPES, "kerneltype not implemented"
    if kerneltype == "diffusion":
        assert th[0] < 1, "Lambda must be < 1"
        if expm_mode == "sparse":      # the literal 2018 semantics (SURVEY A-1)
            K = scipy.sparse.linalg.expm(scipy.sparse.csc_matrix(-th[0] * Lnorm)).toarray()
        elif expm_mode == "dense":     # modern dense path (same algorithm family, dense LU solve)
            K = scipy.linalg.expm(-th[0] * Lnorm)
        elif expm_mode == "spectral":  # independent cross-check: K = V exp(-beta lam) V^T
            lam, V = np.linalg.eigh(Lnorm)
            K = (V * np.exp(-th[0] * lam)) @ V.T
        else:
            raise ValueError(expm_mode)
    elif kerneltype == "regularized_laplacian":
        if expm_mode == "spectral":
            lam, V = np.linalg.eigh(Lnorm)
            K = (V / (1.0 + th[0] * lam)) @ V.T
        else:
            K = scipy.linalg.inv(np.eye(N) + th[0] * Lnorm)
    else:
        assert th[0] >= 2, "a must be >=2"
        K = np.asarray(np.linalg.matrix_power(th[0] * np.eye(N) - Lnorm, pstep_order(th)))
    return np.asarray(K)


def kernel_block(Kfull, rows, cols, theta, measnoise=1.0):
    """``K[sorted rows][:, sorted cols] + measnoise*exp(theta[-1])`` (GPnet.py:305-323,362-365).
    ``rows``/``cols`` are graph POSITIONS; output is in ascending position order (quirk A-5)."""
    th = np.atleast_1d(np.exp(np.squeeze(np.asarray(theta, dtype=np.float64))))
    r = np.array(sorted(set(int(x) for x in rows)), dtype=np.int64)
    c = np.array(sorted(set(int(x) for x in cols)), dtype=np.int64)
    return Kfull[np.ix_(r, c)] + measnoise * th[-1]


def kernel_derivatives(Lnorm, Kfull, kerneltype, theta):
    """d K_full / d theta_0 for the three kernels (NOT in the reference -- SURVEY Appendix C):
    diffusion -beta L K; regularized Laplacian K^2 - K; p-step a p B^(p-1)."""
    th = np.atleast_1d(np.exp(np.squeeze(np.asarray(theta, dtype=np.float64))))
    N = Lnorm.shape[0]
    if kerneltype == "diffusion":
        return -th[0] * (Lnorm @ Kfull)
    if kerneltype == "regularized_laplacian":
        return Kfull @ Kfull - Kfull
    p = pstep_order(th)
    B = th[0] * np.eye(N) - Lnorm
    return th[0] * p * np.asarray(np.linalg.matrix_power(B, p - 1)) if p >= 1 else np.zeros_like(Kfull)


# ----------------------------------------------------------------------------------------------
# GP regression                                                         (GPnetRegressor.py:79-150)
# ----------------------------------------------------------------------------------------------
def gpr_neg_logp(Ky, t):
    """``logPosterior`` body after the kernel call (GPnetRegressor.py:140-150).  Returns ``-logp``;
    ``-inf`` if Cholesky fails (quirk A-8).  Triangular solves via general ``solve`` as the reference."""
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    try:
        L = np.linalg.cholesky(Ky)
    except np.linalg.LinAlgError:
        return -np.inf
    alpha = np.linalg.solve(L.T, np.linalg.solve(L, t))
    logp = -0.5 * np.dot(t, alpha) - np.sum(np.log(np.diag(L))) - Ky.shape[0] * 0.5 * np.log(2 * np.pi)
    return float(-logp)


def gpr_neg_logp_grad(Ky, dKs, t):
    """Analytic gradient of ``-logp`` following old_implementation/GPR.py:52-64 (== GPRnet.py:19-34,
    scripts/prova.py:67-79): ``dlogp/dtheta_j = 1/2 t^T K^-1 dK_j K^-1 t - 1/2 tr(K^-1 dK_j)``;
    the function returns the NEGATIVE gradient like the reference's ``gradLogPosterior``."""
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    L = np.linalg.cholesky(Ky)
    n = Ky.shape[0]
    invk = scipy.linalg.cho_solve((L, True), np.eye(n))
    a = invk @ t
    g = np.zeros(len(dKs))
    for j, dK in enumerate(dKs):
        g[j] = 0.5 * a @ (dK @ a) - 0.5 * np.sum(invk * dK)
    return -g


def gpr_predict(k, kstar, kstarstar, t):
    """GPnetRegressor.predict body, GPnetRegressor.py:86-128 (without the eigvals PD gates)."""
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    t_mean = np.mean(t)
    t_shifted = t - t_mean
    L = np.linalg.cholesky(k)
    alpha = np.linalg.solve(L.T, np.linalg.solve(L, t_shifted))
    fstar = kstar @ alpha + t_mean
    v = np.linalg.solve(L, kstar.T)
    kss_diag = np.diag(kstarstar)
    Vdiag = kss_diag - np.sum(v * v, axis=0)        # == diag(kstarstar_diag - v^T v), :127-128
    s = np.sqrt(Vdiag)
    return dict(L=L, alpha=alpha, fstar=fstar, v=v, Vdiag=Vdiag, s=s, t_mean=t_mean)


# ----------------------------------------------------------------------------------------------
# GP classification, Laplace approximation                       (GPnetClassifier.py:90-147,184-240)
# ----------------------------------------------------------------------------------------------
def gpc_newton(K, y, tol=0.1, phif=1e100, scale=1.0, logq_variant="current", max_outer=100000):
    """``NRiteration`` body after the kernel call, GPnetClassifier.py:95-147.  ``K`` already contains
    the all-entries noise.  Returns ``(f, logq, a, iterations)``.

    ``logq_variant="current"`` is GPnetClassifier.py:138-145 (quirk A-10); ``"notebook"`` is the older
    expression of scripts/prova.py:197-200 == report_notebook.ipynb cell 36, the one the reference's
    only known-answer value (cell 49) pins."""
    n = K.shape[0]
    y = np.asarray(y, dtype=np.float64).reshape(n, 1)
    f = np.zeros((n, 1))
    count = 0
    iters = 0
    Kinv = np.linalg.inv(K)          # the reference recomputes this every iteration (:126); K is fixed
    while True:
        count += 1
        iters += 1
        w = np.squeeze(np.exp(-f) / (1 + np.exp(-f)) ** 2, axis=1)       # :106
        sw = np.sqrt(w)
        B = np.eye(n) + sw[:, None] * K * sw[None, :]                    # :110 (diag products)
        L = np.linalg.cholesky(B)
        p = 1 / (1 + np.exp(-f))                                         # :111
        b = w[:, None] * f + 0.5 * (y + 1) - p                           # :112
        rhs = sw[:, None] * (K @ b)
        a = scale * (b - sw[:, None] * np.linalg.solve(L.T, np.linalg.solve(L, rhs)))   # :113-121
        f = K @ a                                                        # :122
        oldphif = phif
        phif = (np.log(p) - 0.5 * (f.T @ (Kinv @ f)) - 0.5 * np.sum(np.log(np.diag(L)))
                - n / 2.0 * np.log(2 * np.pi))                           # :124-129 (n-vector)
        if np.sum((oldphif - phif) ** 2) < tol:                          # :132
            break
        elif count > 100:                                                # :134-136
            count = 0
            scale = scale / 2.0
        if iters >= max_outer:
            raise RuntimeError("Newton loop did not terminate")
    s = -y * f
    if logq_variant == "current":
        lik = np.sum(np.log(1 + np.log(1 + np.exp(-s))))                 # :140-143 (sic)
    else:
        ps = np.where(s > 0, s, 0)
        lik = np.sum(np.log(ps + np.log(np.exp(-ps) + np.exp(s - ps))))
    logq = -0.5 * float((a.T @ f).item()) - lik - np.sum(np.log(np.diag(L)))
    return f, float(logq), a, iters


def gpc_predict(K, kstar, kss_diag, y, f):
    """GPnetClassifier.predict after NRiteration, GPnetClassifier.py:203-233.  ``kstar`` is n x m
    (train x test, no noise), ``kss_diag`` the diagonal of K[test,test] without noise."""
    n = K.shape[0]
    y = np.asarray(y, dtype=np.float64).reshape(n, 1)
    s = np.where(f < 0, f, 0)
    w = np.squeeze(np.exp(2 * s - f) / ((np.exp(s) + np.exp(s - f)) ** 2), axis=1)
    sw = np.sqrt(w)
    L = np.linalg.cholesky(np.eye(n) + sw[:, None] * K * sw[None, :])
    p = np.exp(s) / (np.exp(s) + np.exp(s - f))
    fstar = np.squeeze(kstar.T @ ((y + 1) * 0.5 - p))
    v = np.linalg.solve(L, sw[:, None] * kstar)
    V = kss_diag - np.sum(v * v, axis=0)
    pi_star = pi_star_from(fstar, V)
    probs = np.vstack((1 - pi_star, pi_star)).T
    return fstar, V, probs


def gpc_grad(K, dKs, y, f, a):
    """GPnetClassifier.gradLogPosterior body after the kernel / NRiteration calls, GPnetClassifier.py:156-182 (RW Alg. 5.1
    as written there; the reference cannot execute it today because its kernel() no longer returns derivative slices,
    quirk A-12).  ``K``: train block incl. noise, ``dKs``: list of dK/dtheta_d blocks, ``f``/``a``: Newton outputs.
    Returns ``-gradZ``."""
    n = K.shape[0]
    y = np.asarray(y, dtype=np.float64).reshape(n, 1)
    f = np.asarray(f, dtype=np.float64).reshape(n, 1)
    a = np.asarray(a, dtype=np.float64).reshape(n, 1)
    s = np.where(f < 0, f, 0)
    w = np.squeeze(np.exp(2 * s - f) / ((np.exp(s) + np.exp(s - f)) ** 2), axis=1)
    sw = np.sqrt(w)
    L = np.linalg.cholesky(np.eye(n) + sw[:, None] * K * sw[None, :])                      # :159
    R = sw[:, None] * np.linalg.solve(L.T, np.linalg.solve(L, np.diag(sw)))                # :161
    C = np.linalg.solve(L, sw[:, None] * K)                                                # :162
    p = np.exp(s) / (np.exp(s) + np.exp(s - f))
    hess = -np.exp(2 * s - f) / (np.exp(s) + np.exp(s - f)) ** 2
    s2 = -0.5 * (np.diag(K) - np.sum(C * C, axis=0))[:, None] * (2 * hess * (0.5 - p))     # :165-168
    gradZ = np.zeros(len(dKs))
    for d, dK in enumerate(dKs):
        s1 = 0.5 * float((a.T @ (dK @ a)).item()) - 0.5 * np.trace(R @ dK)                 # :174-176
        b = dK @ ((y + 1) * 0.5 - p)                                                       # :177
        s3 = b - K @ (R @ b)                                                               # :179
        gradZ[d] = s1 + float((s2.T @ s3).item())
    return -gradZ


def pi_star_from(fstar, V):
    """5-erf averaged sigmoid, GPnetClassifier.py:221-231 (same expression order)."""
    m = len(V)
    alpha = np.tile((1 / (2 * V)), (5, 1))
    gamma = np.einsum("i,k->ik", LAMBDAS, fstar.T)
    lambdas_mat = np.tile(LAMBDAS, (m, 1)).T
    Vmat = np.tile(V, (5, 1))
    integrals = (np.sqrt(np.pi / alpha) * erf(gamma * np.sqrt(alpha / (alpha + lambdas_mat ** 2)))
                 / (2 * np.sqrt(Vmat * 2 * np.pi)))
    return (COEFS * integrals).sum(axis=0) + 0.5 * COEFS.sum()


def predicted_class(probs):
    """GPnet.py:587-603: ``-( +1 if prob_0 > 0.5 else -1 )``."""
    p0 = probs[:, 0]
    return -np.where(p0 > 0.5, 1.0, -1.0)


# ----------------------------------------------------------------------------------------------
# a tiny model object mirroring the reference call pattern (kernel rebuilt per call, A-17)
# ----------------------------------------------------------------------------------------------
class OracleGP:
    """Holds a graph's normalized Laplacian; mirrors ``GPnetBase.kernel`` + the model methods on
    graph POSITIONS.  ``expm_mode`` selects the diffusion flavour (see ``full_kernel``)."""

    def __init__(self, G=None, csr=None, kerneltype="diffusion", expm_mode="sparse"):
        if csr is None:
            csr = graph_to_csr(G)[:3]
        self.indptr, self.indices, self.weights = csr
        self.N = len(self.indptr) - 1
        self.kerneltype = kerneltype
        self.expm_mode = expm_mode
        self.Lnorm = normalized_laplacian(self.indptr, self.indices, self.weights, self.N)

    def full(self, theta):
        return full_kernel(self.Lnorm, self.kerneltype, theta, self.expm_mode)

    def kernel(self, rows, cols, theta, measnoise=1.0, Kfull=None):
        Kfull = self.full(theta) if Kfull is None else Kfull
        return kernel_block(Kfull, rows, cols, theta, measnoise)

    # regressor ------------------------------------------------------------------------------
    def neg_logp(self, theta, train, t):
        return gpr_neg_logp(self.kernel(train, train, theta), t)

    def neg_logp_grad(self, theta, train, t):
        """(-logp, -grad) with the analytic gradient; P = len(theta)."""
        theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
        P = len(theta)
        Kfull = self.full(theta)
        Ky = kernel_block(Kfull, train, train, theta)
        r = np.array(sorted(train), dtype=np.int64)
        dK0 = kernel_derivatives(self.Lnorm, Kfull, self.kerneltype, theta)[np.ix_(r, r)]
        c = float(np.exp(theta[-1]))
        dKs = []
        for j in range(P):
            D = np.zeros_like(Ky)
            if j == 0:
                D = D + dK0
            if j == P - 1:
                D = D + c
            dKs.append(D)
        return gpr_neg_logp(Ky, t), gpr_neg_logp_grad(Ky, dKs, t)

    def predict_regressor(self, theta, train, test, t):
        Kfull = self.full(theta)
        k = kernel_block(Kfull, train, train, theta)
        kstar = kernel_block(Kfull, test, train, theta, measnoise=0.0)
        kss = kernel_block(Kfull, test, test, theta)
        out = gpr_predict(k, kstar, kss, t)
        out.update(k=k, kstar=kstar, kstarstar_diag=np.diag(kss).copy())
        return out

    def landscape(self, theta, axidx, ax1, ax2, train, t):
        """GPnet.py:539-552."""
        lml = np.zeros([len(ax1), len(ax2)])
        params = list(theta)
        for i in range(len(ax1)):
            for j in range(len(ax2)):
                params[axidx[0]] = ax1[i]
                params[axidx[1]] = ax2[j]
                lml[i, j] = -self.neg_logp(params, train, t)
        return lml

    # classifier -----------------------------------------------------------------------------
    def newton(self, theta, train, y, **kw):
        return gpc_newton(self.kernel(train, train, theta), y, **kw)

    def classifier_grad(self, theta, train, y):
        """(-logq, -gradZ): NRiteration + gradLogPosterior with the derivative kernels of SURVEY Appendix C."""
        theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
        P = len(theta)
        Kfull = self.full(theta)
        K = kernel_block(Kfull, train, train, theta)
        r = np.array(sorted(train), dtype=np.int64)
        dK0 = kernel_derivatives(self.Lnorm, Kfull, self.kerneltype, theta)[np.ix_(r, r)]
        c = float(np.exp(theta[-1]))
        dKs = []
        for j in range(P):
            D = np.zeros_like(K)
            if j == 0:
                D = D + dK0
            if j == P - 1:
                D = D + c
            dKs.append(D)
        f, logq, a, iters = gpc_newton(K, y)
        return -logq, gpc_grad(K, dKs, y, f, a)

    def predict_classifier(self, theta, train, test, y):
        Kfull = self.full(theta)
        K = kernel_block(Kfull, train, train, theta)
        kstar = kernel_block(Kfull, train, test, theta, measnoise=0.0)
        kss_diag = np.diag(kernel_block(Kfull, test, test, theta, measnoise=0.0)).copy()
        f, logq, a, iters = gpc_newton(K, y)
        fstar, V, probs = gpc_predict(K, kstar, kss_diag, y, f)
        return dict(f=f, logq=logq, a=a, iters=iters, fstar=fstar, V=V, probs=probs,
                    labels=predicted_class(probs))


# ----------------------------------------------------------------------------------------------
# the reference's only known-answer test                     (report_notebook.ipynb cells 14-49)
# ----------------------------------------------------------------------------------------------
def notebook_kat_inputs():
    """Inputs of report_notebook.ipynb cells 43,45,48: N=40, RandomState(4), labels cos(0.7x)<0 -> -1,
    theta=[1,0,0.5]; kernel of cell 16 (squared exponential + noise on the DIAGONAL)."""
    rng = np.random.RandomState(4)
    x = np.vstack(rng.uniform(-5, 5, 40))
    y = np.vstack(np.where(np.cos(0.7 * x).flatten() < 0, -1, 1))
    th = np.exp(np.array([1.0, 0.0, 0.5]))
    d2 = (x - x.T) ** 2 * th[1]
    K = th[0] * np.exp(-0.5 * d2) + th[2] * np.eye(40)
    return K, y


NOTEBOOK_KAT_VALUE = -77.22939013   # report_notebook.ipynb cell 49 output (logPosterior = -logq)
