"""CPU tier: the algebra of the spectral path (csrc/spectral.cu) restated in numpy and checked against the oracle, and the host
logic around it.  No GPU, no product code path that could compute on the CPU: these restate what the kernels do."""
import numpy as np
import pytest

from oracle import configs as cfg
from oracle import gp_oracle as go


def round_maps(nb):
    """tests' restatement of round_maps() in spectral.cu: block-level gather permutations of the first and of every later round."""
    arr = list(range(nb))
    out = []
    for _ in range(2):
        src = [0] * nb
        for i in range(nb // 2):
            src[2 * i], src[2 * i + 1] = arr[i], arr[nb - 1 - i]
        out.append(src)
        na = [0] * nb
        for i in range(nb // 2):
            na[i], na[nb - 1 - i] = 2 * i, 2 * i + 1
        arr = [na[0], na[nb - 1]] + na[1:nb - 1]
    return out


@pytest.mark.parametrize("nb", [2, 4, 6, 16, 30])
def test_round_robin_on_physical_slots_meets_every_pair_once_per_sweep(nb):
    first, later = round_maps(nb)
    ids = list(range(nb))
    met = set()
    for rnd in range(nb - 1):
        src = first if rnd == 0 else later
        ids = [ids[s] for s in src]
        for i in range(nb // 2):
            pair = tuple(sorted((ids[2 * i], ids[2 * i + 1])))
            assert pair not in met
            met.add(pair)
    assert len(met) == nb * (nb - 1) // 2


def test_one_jacobi_round_as_the_kernels_apply_it():
    """A' = J^T (P A P^T) J and V'^T = J^T P V^T computed the way spectral.cu does (row gather, Q^T from the left, transposed
    output, row gather again, Q^T from the left) equal the direct formulas; the spectrum is preserved."""
    rng = np.random.RandomState(0)
    b, nb = 4, 6
    N = b * nb
    A = rng.randn(N, N)
    A = A + A.T
    VT = np.linalg.qr(rng.randn(N, N))[0]
    src = round_maps(nb)[0]
    rm = np.concatenate([np.arange(s * b, (s + 1) * b) for s in src])
    P = np.eye(N)[rm]
    J = np.zeros((N, N))
    for p in range(nb // 2):
        sl = slice(2 * b * p, 2 * b * (p + 1))
        S = (P @ A @ P.T)[sl, sl]
        J[sl, sl] = np.linalg.eigh(S)[1]
    AP = A[rm]                                   # gather_rows(A)
    BT = (J.T @ AP).T                            # batched GEMM, only the transposed output kept
    A_new = J.T @ BT[rm]                         # gather_rows(BT), second batched GEMM
    VT_new = J.T @ VT[rm]
    assert np.allclose(A_new, J.T @ P @ A @ P.T @ J, atol=1e-12)
    assert np.allclose(np.sort(np.linalg.eigvalsh((A_new + A_new.T) / 2)), np.sort(np.linalg.eigvalsh(A)), atol=1e-11)
    # V^T rows follow the same permutation: V'^T A_orig-basis relation  A = V diag V^T is carried along
    assert np.allclose(VT_new @ VT_new.T, np.eye(N), atol=1e-12)
    for p in range(nb // 2):                      # the pivot blocks are diagonal afterwards
        sl = slice(2 * b * p, 2 * b * (p + 1))
        blk = A_new[sl, sl]
        assert np.max(np.abs(blk - np.diag(np.diag(blk)))) < 1e-11


def test_rotation_without_division_by_apq_equals_the_textbook_one():
    rng = np.random.RandomState(1)
    for _ in range(200):
        app, aqq, apq = rng.randn(3) * 10.0 ** rng.randint(-8, 3)
        tau = (aqq - app) / (2 * apq)
        t_ref = np.sign(tau) / (abs(tau) + np.sqrt(1 + tau * tau)) if tau != 0 else 1.0
        d, b2 = aqq - app, 2 * apq
        t = b2 / (d + np.copysign(np.hypot(d, b2), d))
        assert abs(t - t_ref) <= 1e-14 * max(abs(t_ref), 1e-300)
        c = 1 / np.sqrt(1 + t * t)
        s = t * c
        Jm = np.array([[c, s], [-s, c]])
        Sm = Jm.T @ np.array([[app, apq], [apq, aqq]]) @ Jm
        assert abs(Sm[0, 1]) <= 1e-13 * (abs(app) + abs(aqq) + abs(apq))


@pytest.mark.parametrize("kt,theta", [("diffusion", np.log([0.6, 0.01])), ("regularized_laplacian", np.log([0.7, 0.05])),
                                      ("pstep_walk", np.log([2.3, 4, 0.01]))])
def test_batched_evaluator_algebra_against_the_oracle(kt, theta):
    """What spectral_batch_kernel computes per theta, restated in numpy: K_tt = V_t r V_t^T + c, D = V_t r' V_t^T, Cholesky,
    W = L^-1, K_y^-1 = W^T W, -logp = q/2 + sum log L_ii + n/2 log 2pi, gradient -(a^T D a - tr(K^-1 D))/2 and
    -c ((sum a)^2 - sum K^-1)/2  --  against OracleGP.neg_logp_grad."""
    G, tr, te, t = cfg.grid_case(12, 12, 40, 60)
    ip, ix, w = go.graph_to_csr(G)[:3]
    N = len(ip) - 1
    lam, V = np.linalg.eigh(go.normalized_laplacian(ip, ix, w, N))
    e = np.exp(theta)
    th0, noise = e[0], e[-1]
    if kt == "diffusion":
        r = np.exp(-th0 * lam); rd = -th0 * lam * r
    elif kt == "regularized_laplacian":
        r = 1 / (1 + th0 * lam); rd = r * r - r
    else:
        p = int(e[1]); r = (th0 - lam) ** p; rd = th0 * p * (th0 - lam) ** (p - 1)
    Vt = V[tr]                                    # n x N (the kernel holds its transpose: row k = eigenvector k on the training nodes)
    Ky = (Vt * r) @ Vt.T + noise
    D = (Vt * rd) @ Vt.T
    L = np.linalg.cholesky(Ky)
    W = np.linalg.inv(L)
    Kinv = W.T @ W
    tt = t[tr]
    a = Kinv @ tt
    n = len(tr)
    val = 0.5 * tt @ a + np.sum(np.log(np.diag(L))) + 0.5 * n * np.log(2 * np.pi)
    g = np.zeros(len(theta))
    g[0] += -0.5 * (a @ D @ a - np.sum(Kinv * D))
    g[-1] += -0.5 * noise * (np.sum(a) ** 2 - np.sum(Kinv))
    oval, ograd = go.OracleGP(G, kerneltype=kt, expm_mode="dense").neg_logp_grad(theta, tr, tt)
    assert abs(val - oval) <= 1e-10 * abs(oval)
    assert np.max(np.abs(g - ograd)) <= 1e-8 * np.max(np.abs(ograd))


def test_spectral_landscape_branch_builds_the_reference_loop_thetas():
    """sharded.evaluate_points on the spectral path hands the batched evaluator exactly the parameter vectors the reference's
    double loop visits (GPnet.py:542-549, second axis written last) and flips the sign / maps a failed Cholesky to +inf."""
    from gpnet_b200 import sharded

    class Ctx:
        def gpr_logp_grad_batch(self, kind, thetas, t, want_grad):
            self.thetas = np.array(thetas)
            out = np.zeros((len(thetas), 1 + thetas.shape[1]))
            out[:, 0] = np.arange(len(thetas)) + 1.0
            out[:, 1:] = 2.0
            info = np.zeros(len(thetas), dtype=np.int32)
            info[1] = 3
            return 0, out, info

    class Engine:
        spectral = True

    class Model:
        kerneltype = "diffusion"
        training_nodes = list(range(5))
        test_nodes = []
        training_values = np.zeros(5)
        _engine = Engine()
        logPosterior_and_grad = True

        def _theta_vec(self, th):
            return np.asarray(th, dtype=float)

        def _positions(self, nodes):
            return list(nodes)

    ctx = Ctx()
    m = Model()
    orig = sharded._bound_context
    sharded._bound_context = lambda model: ctx
    try:
        ax1, ax2 = np.array([0.1, 0.2, 0.3]), np.array([-1.0, -2.0])
        out = sharded.evaluate_points(m, [9.0, 9.0], [0, 1], ax1, ax2, [0, 3, 5], want_grad=False)
    finally:
        sharded._bound_context = orig
    assert np.array_equal(ctx.thetas, np.array([[0.1, -1.0], [0.2, -2.0], [0.3, -2.0]]))
    assert out[0, 0] == -1.0 and np.isposinf(out[1, 0]) and out[2, 0] == -3.0 and np.all(out[:, 1:] == 0.0)


def test_spectral_landscape_branch_of_a_classifier_uses_the_batched_newton_loop():
    """The classifier's landscape on the spectral path hands the same loop thetas to gpn_gpc_newton_batch and maps a failed
    Cholesky inside the loop to NaN (like gpn_landscape_points)."""
    from gpnet_b200 import sharded

    class Ctx:
        def gpc_newton_batch(self, kind, thetas, y):
            self.thetas, self.y = np.array(thetas), np.array(y)
            B = len(thetas)
            info = np.zeros(B, dtype=np.int32)
            info[2] = 5
            return 0, -np.arange(1.0, B + 1.0), np.full(B, 4, dtype=np.int32), info

    class Engine:
        spectral = True

    class Model:                                  # no logPosterior_and_grad attribute: a classifier
        kerneltype = "regularized_laplacian"
        training_nodes = list(range(6))
        test_nodes = []
        training_values = np.array([1, -1, 1, 1, -1, -1])
        _engine = Engine()

        def _theta_vec(self, th):
            return np.asarray(th, dtype=float)

        def _positions(self, nodes):
            return list(nodes)

    ctx = Ctx()
    orig = sharded._bound_context
    sharded._bound_context = lambda model: ctx
    try:
        out = sharded.evaluate_points(Model(), [0.0, 0.0], [1, 0], np.array([0.5, 0.6]), np.array([-1.0, -2.0, -3.0]), [0, 1, 5])
    finally:
        sharded._bound_context = orig
    # axidx = [1, 0]: the first axis writes entry 1, the second entry 0
    assert np.array_equal(ctx.thetas, np.array([[-1.0, 0.5], [-2.0, 0.5], [-3.0, 0.6]]))
    assert np.array_equal(ctx.y, Model.training_values.astype(float))
    assert out[0, 0] == -1.0 and out[1, 0] == -2.0 and np.isnan(out[2, 0]) and out.shape == (3, 3)
