"""Drop-in ``GPnetClassifier`` (reference: GPnetClassifier.py:9-240) on the sm_100a library.

The Laplace mode-finding Newton loop (``NRiteration``), ``logPosterior`` and ``predict`` keep the reference's
signatures, shapes and quirks: the log-of-log ``logq`` (A-10), the n-vector convergence test with ``tol = 0.1``
(A-11), the 5-erf averaged sigmoid, and the label rule of ``generate_df`` (GPnet.py:587-603).
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from ._lib import KERNEL_IDS
from .engine import invalidate_predict, raise_domain
from .GPnet import COEFS, LAMBDAS, GPnetBase, KERNEL_LIST  # noqa: F401  (re-exported like the reference module)


class GPnetClassifier(GPnetBase):
    def __init__(self, Graph=False, totnodes=False, ntrain=100, ntest=90, deg=4, seed=0, training_nodes=False,
                 training_values=False, test_nodes=False, theta=[0.1, 0.1, 0.1], optimize=False, relabel_nodes=False,
                 kerneltype="diffusion"):
        super().__init__(Graph, totnodes, ntrain, ntest, deg, seed, training_nodes, training_values, test_nodes, theta,
                         optimize, relabel_nodes, kerneltype)
        self.pivot_flag = False
        if not isinstance(training_values, pd.Series):
            print("no training labels where specified")
            print("> Setting labels to (np.sin(0.6 * self.pvtdist) > 0).replace({True: 1, False: -1})")
            self.pivot_flag = True
            self.pvtdist = self.pivot_distance(0)
            self.binary_labels = pd.Series(np.where(np.sin(0.6 * self.pvtdist) > 0, 1, -1), index=self.pvtdist.index)
            self.training_values = self.binary_labels[self.training_nodes]
        else:
            self.training_values = training_values

    def _kind(self):
        assert self.kerneltype in KERNEL_LIST, "kerneltype not implemented"
        return KERNEL_IDS[self.kerneltype]

    @staticmethod
    def _labels(targets):
        return np.asarray(targets.values if isinstance(targets, (pd.Series, pd.DataFrame)) else targets,
                          dtype=np.float64).reshape(-1)

    # ---- GPnetClassifier.py:85-147 ---------------------------------------------------------------------
    def logPosterior(self, theta, *args):
        data, targets = args
        (f, logq, a) = self.NRiteration(data, targets, theta)
        return -logq

    def NRiteration(self, data, targets, theta, tol=0.1, phif=1e100, scale=1.0):
        th = self._theta_vec(theta)
        ctx = self._engine.bind(self._positions(data), self._positions(self.test_nodes))
        ctx._kernel_key = None
        invalidate_predict(ctx)
        st, r = ctx.gpc_newton(self._kind(), th, self._labels(targets), tol=tol, phif=phif, scale=scale, variant=0)
        raise_domain(st)
        if r["info"] > 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite")   # np.linalg.cholesky, GPnetClassifier.py:110
        self.newton_iterations = r["iters"]
        n = len(r["f"])
        return (r["f"].reshape(n, 1), np.array([[r["logq"]]]), r["a"].reshape(n, 1))

    # ---- GPnetClassifier.py:184-240 --------------------------------------------------------------------
    def predict(self):
        if self.optimize != False:   # noqa: E712
            self.optimize_params()
        th = self._theta_vec(self.theta)
        ctx = self._engine.bind(self._positions(self.training_nodes), self._positions(self.test_nodes))
        ctx._kernel_key = None
        invalidate_predict(ctx)
        st, r = ctx.gpc_predict(self._kind(), th, self._labels(self.training_values))
        raise_domain(st)
        if r["info"] > 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        self.newton_iterations = r["iters"]
        self.fstar = r["fstar"]
        self.V = r["V"]
        pi_star = r["pi_star"]
        self.predicted_probs = np.vstack((1 - pi_star, pi_star)).T
        self.s = np.sqrt(self.V)
        print("succesfully trained model")
        self.is_trained = True
        self._defer_df()       # the reference's generate_df() (GPnet.py:576-605), materialised on first access of self.df
        return (self.fstar.T, self.V, self.predicted_probs)

    def gradLogPosterior(self, theta, *args):
        """RW Alg. 5.1 exactly as the reference writes it (GPnetClassifier.py:149-182), fed by the derivative kernels the
        reference's own ``kernel`` no longer provides (it raises IndexError there, SURVEY A-12).  Returns ``-gradZ``.  Note that
        this is the gradient of the true Laplace LML, not of the reference's ``logq`` (A-10), so ``optimize_params`` stays
        function-only like the reference's."""
        data, targets = args
        th = self._theta_vec(theta)
        ctx = self._engine.bind(self._positions(data), self._positions(self.test_nodes))
        ctx._kernel_key = None
        invalidate_predict(ctx)
        st, out, iters, info = ctx.gpc_grad(self._kind(), th, self._labels(targets))
        raise_domain(st)
        if info > 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        self.newton_iterations = iters
        return out[1:].copy()

    # ---- plotting pass-through (presentation only) ---------------------------------------------------
    def gen_cmap(self):
        """Probabilities live in [0, 1] (GPnetClassifier.py:431-434)."""
        self.vmin, self.vmax = 0, 1
        pl = self._pyplot()
        self.cmap = pl.cm.seismic if pl is not None else None

    def plot_latent(self, filename=False):
        """Latent mean and its +-sqrt(V) band over the pivot distance (GPnetClassifier.py:366-391)."""
        pl = self._pyplot()
        if pl is None:
            return self
        pl.figure(figsize=[15, 9])
        pl.plot(self.test_nodes, self.fstar, "b-")
        pl.gca().fill_between(self.test_nodes, self.fstar - self.s, self.fstar + self.s, color="#dddddd")
        pl.title("Latent Process Mean and Variance")
        self._save(pl, filename)
        return self

    plot_latent_old = plot_latent
    plot_predict_2dold = plot_latent

    def plot_predict_graph(self, filename=False):
        if not self.is_trained:
            print("need to train a model first, use GPnetClassifier.predict()")
            return
        y = (np.asarray([self.training_values[x] for x in self.training_nodes], dtype=np.float64) + 1.0) / 2.0
        return self._plot_on_graph("Prediction plot", y, np.asarray(self.df["prob_1"][list(self.test_nodes)], dtype=np.float64), filename)

    def plot_binary_prediction(self, filename=False):
        """Predicted class and training labels against the pivot distance (GPnetClassifier.py:348-364)."""
        pl = self._pyplot()
        if pl is None:
            return self
        pl.figure(figsize=[15, 9])
        frame = self.df.sort_values(by=["pvtdist"])
        pl.plot(frame["pvtdist"].values, frame["predicted_class"].values, "bo", ms=8)
        pl.plot(frame["pvtdist"].values, frame["train_vals"].values, "r+", ms=20)
        pl.title("Binary Classification")
        pl.xlabel("pivot distance")
        pl.ylabel("classification label")
        pl.legend(("test nodes", "training nodes"), loc="center right")
        self._save(pl, filename)
        return self
