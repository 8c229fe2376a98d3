"""Drop-in ``GPnetBase`` for the dense GP hot path, backed by the sm_100a library (no CPU fallback).

Mirrors the public surface of the reference base class (``GPnet.py:66-609`` in filippocastelli/GPnet):
constructor arguments, node bookkeeping, ``kernel``, ``optimize_params``, ``lml_landscape``, ``logp``,
``calc_ktot``, ``is_pos_def``, ``generate_df``, ``set_training_values``, ``pivot_distance``.  The
numerics run on the GPU through ``gpnet_b200._lib``; only graph bookkeeping, the scipy optimiser and
the pandas result frame stay on the host.  Behavioural differences are limited to the fixes listed in
SURVEY.md A-15 (python >= 3.11 sampling, Series-valued ``training_values``) and to laziness: the
all-pairs shortest paths and the Kamada-Kawai layout (``GPnet.py:204-207``) are computed on first
access of ``dist`` / ``plot_pos`` instead of in ``__init__``.
"""
from __future__ import annotations

import random

import networkx as nx
import numpy as np
import pandas as pd
import scipy.optimize as so

from .engine import GraphEngine

# GPnet.py:59-62 -- constants of the 5-erf approximation of the logistic sigmoid (also baked into the
# device kernel pistar_kernel in csrc/gp.cu)
LAMBDAS = np.array([0.41, 0.4, 0.37, 0.44, 0.39])
COEFS = np.array([-1854.8214151, 3516.89893646, 221.29346712, 128.12323805, -2010.49422654])[:, np.newaxis]

KERNEL_LIST = ("diffusion", "regularized_laplacian", "pstep_walk")


def _is_false(x) -> bool:
    """``x == False`` of the reference, safe for lists / Series (SURVEY A-15)."""
    return isinstance(x, (bool, np.bool_)) and not x


class GPnetBase:
    """Common state and drivers of ``GPnetRegressor`` / ``GPnetClassifier`` (reference: GPnet.py:66)."""

    def __init__(self, Graph, totnodes, ntrain, ntest, deg, seed, training_nodes, training_values, test_nodes, theta,
                 optimize, relabel_nodes, kerneltype):
        self.N = ntrain
        self.n = ntest
        self.deg = deg
        self.seed = seed
        self.is_trained = False
        self.optimize = optimize
        self.theta = theta
        self.relabel_nodes = relabel_nodes
        self.kerneltype = kerneltype
        self.totnodes = (self.N + self.n) if _is_false(totnodes) else totnodes

        if _is_false(Graph):
            print("> Initializing Random Regular Graph")
            print(self.totnodes, "nodes")
            print("node degree", self.deg)
            G = nx.random_regular_graph(deg, self.totnodes)
        else:
            G = Graph
            self.totnodes = len(Graph.nodes)
        self.Graph = G

        self.orig_labels_dict = dict(zip(G, range(len(G.nodes))))
        self.orig_labels_invdict = {v: k for k, v in self.orig_labels_dict.items()}
        if relabel_nodes == True:   # noqa: E712  (reference accepts any truthy-equal value)
            print("> Relabeling nodes, orig. names stored in self.orig_labels_dict")
            self.Graph = nx.relabel_nodes(G, self.orig_labels_dict)

        self._engine = GraphEngine(self.Graph)
        self._dist = None
        self._plot_pos = None

        self.training_nodes = training_nodes
        self.test_nodes = test_nodes
        if _is_false(training_nodes) or _is_false(test_nodes):
            print("> Assigning Nodes Randomly ( seed =", self.seed, ")")
            print(self.N, " training nodes")
            print(self.n, " test nodes")
            print((self.totnodes - (self.N + self.n)), " idle nodes")
            self.random_assign_nodes()
        self.assign_other_nodes()
        self.calc_pivot_distance()

    # ---- lazily evaluated, hot-path-irrelevant attributes (GPnet.py:204-207) -----------------------
    @property
    def dist(self):
        if self._dist is None:
            self.calc_shortest_paths()
        return self._dist

    @property
    def plot_pos(self):
        if self._plot_pos is None:
            self._plot_pos = nx.kamada_kawai_layout(self.Graph)
        return self._plot_pos

    def calc_shortest_paths(self):
        lengths = dict(nx.all_pairs_shortest_path_length(self.Graph))
        self._dist = pd.DataFrame(lengths).sort_index(axis=1)

    def pivot_distance(self, pivot=0):
        d = pd.Series(dict(nx.single_source_shortest_path_length(self.Graph, pivot))).sort_index()
        d.name = "pivot distance"
        return d

    def calc_pivot_distance(self):
        self.pvtdist = self.pivot_distance(list(self.Graph.nodes)[0])
        return self

    # ---- node assignment (GPnet.py:229-257) ------------------------------------------------------
    def random_assign_nodes(self):
        if self.N + self.n > self.totnodes:
            raise ValueError("tot. nodes cannot be less than training nodes + test nodes")
        random.seed(self.seed)
        nodes = list(self.Graph.nodes)
        self.training_nodes = sorted(random.sample(nodes, self.N))
        chosen = set(self.training_nodes)
        rest = [x for x in nodes if x not in chosen]      # ordered population (python >= 3.11 forbids sets)
        self.test_nodes = sorted(random.sample(rest, self.n))
        self.assign_other_nodes()
        return self

    def assign_other_nodes(self):
        others = set(self.Graph.nodes) - set(self.training_nodes) - set(self.test_nodes)
        self.other_nodes = sorted(others)
        return self

    # ---- kernel (GPnet.py:276-367) -----------------------------------------------------------------
    def _theta_vec(self, theta):
        th = np.asarray(theta, dtype=np.float64)
        if th.ndim == 0 or len(th) != 1:
            th = np.squeeze(th)
        return np.atleast_1d(th)

    def kernel(self, nodes_a, nodes_b, theta, measnoise=1.0, wantderiv=True):
        """``K[nodes_a, nodes_b] + measnoise*exp(theta[-1])`` as a fresh host array, rows/columns in graph
        order (quirks A-3, A-5).  ``wantderiv`` is accepted and ignored like in the reference."""
        th = self._theta_vec(theta)
        assert self.kerneltype in KERNEL_LIST, "kerneltype not implemented"
        relabelled = self.relabel_nodes == True   # noqa: E712
        rows = self._engine.positions(nodes_a, relabelled)
        cols = self._engine.positions(nodes_b, relabelled)
        ctx = self._engine.build_kernel(self.kerneltype, th, want_grad=False)
        return ctx.kernel_block(rows, cols, float(measnoise) * float(np.exp(th[-1])))

    def _positions(self, nodes):
        return self._engine.positions(nodes, self.relabel_nodes == True)   # noqa: E712

    def logPosterior(self, theta, *args):
        raise NotImplementedError("logPosterior() must be overridden by GPnetRegressor or GPnetClassifier")

    def logp(self):
        return -self.logPosterior(self.theta, self.training_nodes, self.training_values)

    # ---- drivers (GPnet.py:259-274, 539-552) -------------------------------------------------------
    def optimize_params(self, gtol=1e-3, maxiter=200, disp=1):
        """scipy.optimize.minimize over the log-parameters, exactly the call of GPnet.py:264-271.  Extension:
        ``optimize={"method":..., "bounds":..., "jac": True}`` feeds the fused device logp+gradient to the
        optimiser instead of letting scipy take (P+1) finite-difference probes per gradient."""
        if self.optimize != False:   # noqa: E712
            print("> Optimizing parameters")
            print("method used: ", self.optimize["method"])
            print("bounds: ", self.optimize["bounds"])
            use_jac = bool(self.optimize.get("jac", False)) and hasattr(self, "logPosterior_and_grad")
            res = so.minimize(
                fun=self.logPosterior_and_grad if use_jac else self.logPosterior,
                x0=self.theta,
                args=(self.training_nodes, self.training_values),
                method=self.optimize["method"],
                bounds=self.optimize["bounds"],
                jac=True if use_jac else None,
                options={"disp": True},
            )
            self.theta = res["x"]
            self.optimize_result = res
            print("new parameters: ", self.theta)
        return self

    def use_spectral(self, on=True):
        """Extension: build every kernel of this model as ``V r(lam) V^T`` from ONE eigendecomposition of the
        normalized Laplacian (computed on the GPU the first time a kernel is needed; eq. 55-56 of the reference's write-up)
        instead of expm / inv / matrix_power per theta (GPnet.py:338-349).  Pays for sweeps over many kernel parameters
        (``lml_landscape``, ``optimize_params``); results agree with the default path to ~1e-12.  With a training set of at
        most 112 nodes the regressor's landscapes and batches then run one CTA per theta in a single launch."""
        self._engine.spectral = bool(on)
        self._engine.invalidate_kernel()
        return self

    def lml_landscape(self, theta, axidx, ax1, ax2):
        """``lml[i, j] = -logPosterior(theta with theta[axidx[0]] = ax1[i], theta[axidx[1]] = ax2[j])``.
        Under torch.distributed with world_size > 1 the grid is sharded cyclically over the ranks and the
        results are all-gathered (see gpnet_b200.sharded); every rank returns the full array."""
        print("> Calculating LML Landscape")
        from .sharded import sharded_landscape
        lml = sharded_landscape(self, theta, axidx, ax1, ax2)[..., 0]
        try:    # the reference aliases and mutates the caller's theta (A-14): leave the same final state
            theta[axidx[0]] = ax1[len(ax1) - 1]
            theta[axidx[1]] = ax2[len(ax2) - 1]
        except TypeError:
            pass
        return lml

    def plot_lml_landscape(self, plots, params, filename=False):
        """Computes every landscape of ``plots`` (GPnet.py:469-537) and draws them when matplotlib is present."""
        results = {}
        for name, plot in plots.items():
            lml = self.lml_landscape(params, plot[0], plot[1], plot[2])
            idxmax = np.unravel_index(np.argmax(lml, axis=None), lml.shape)
            print(idxmax, lml[idxmax])
            results[name] = lml
        self.lml_landscapes = results
        try:
            import matplotlib.pyplot as pl
        except ImportError:
            return results
        fig, ax = pl.subplots(1, max(len(plots), 1), squeeze=False, dpi=150)
        fig.suptitle("LML landscapes", size=10)
        for idx, (name, plot) in enumerate(plots.items()):
            cax = ax[0][idx].pcolor(plot[2], plot[1], results[name])
            ax[0][idx].set(xlabel="theta" + str(plot[0][0]), ylabel="theta" + str(plot[0][1]), title=name)
            fig.colorbar(cax, ax=ax[0][idx])
        if isinstance(filename, str):
            pl.savefig(filename, bbox_inches="tight")
        return results

    # ---- small utilities (GPnet.py:554-605) ------------------------------------------------------
    def set_training_values(self, training_values):
        self.training_values = training_values
        self.training_values.name = "training values"

    def calc_ktot(self):
        nodes = list(self.Graph.nodes)
        self.ktot = self.kernel(nodes_a=nodes, nodes_b=nodes, theta=self.theta, wantderiv=False)

    def is_pos_def(self, test_mat):
        """PD test through the device Cholesky (the reference uses np.linalg.eigvals, GPnet.py:572-573;
        PD <=> Cholesky succeeds for the symmetric matrices it is applied to, quirk A-7)."""
        import torch
        a = np.ascontiguousarray(np.asarray(test_mat, dtype=np.float64))
        n = a.shape[0]
        ctx = self._engine.ctx
        ld = (n + 15) // 16 * 16
        buf = torch.zeros((n, ld), dtype=torch.float64, device=f"cuda:{ctx.device}")
        buf[:, :n] = torch.from_numpy(a).to(buf.device)
        torch.cuda.synchronize()
        info = ctx.potrf(n, buf.data_ptr(), ld)
        self._engine.invalidate_kernel()
        return info == 0

    # ``df`` (GPnet.py:576-605) is assembled on first access after a predict(): building the pandas frame (and the pivot
    # distances it needs) costs more host time than the whole device side of a C1 / C2 prediction (3-6 ms against 1.5-7 ms), and
    # most callers of predict() in an optimisation loop never look at it.  generate_df() itself stays eager and public.
    @property
    def df(self):
        if self.__dict__.get("_df_pending"):
            self.__dict__["_df_pending"] = False
            self.generate_df()
        return self.__dict__.get("_df")

    @df.setter
    def df(self, value):
        self.__dict__["_df"] = value
        self.__dict__["_df_pending"] = False

    def _defer_df(self):
        """predict() calls this where the reference calls generate_df()."""
        self.__dict__["_df_pending"] = True

    def generate_df(self):
        fstar_series = pd.Series(index=self.test_nodes, data=self.fstar)
        s_series = pd.Series(index=self.test_nodes, data=self.s)
        probs = getattr(self, "predicted_probs", None)
        if probs is None:
            probs0 = pd.Series(index=self.test_nodes, dtype=np.float64)
            probs1 = pd.Series(index=self.test_nodes, dtype=np.float64)
        else:
            probs0 = pd.Series(index=self.test_nodes, data=probs.T[0])
            probs1 = pd.Series(index=self.test_nodes, data=probs.T[1])
        cls = probs0.copy()
        hi = cls > 0.5
        lo = cls <= 0.5
        cls[hi] = 1
        cls[lo] = -1
        self.df = pd.DataFrame().assign(
            pvtdist=self.pvtdist,
            train_vals=self.training_values,
            fstar=fstar_series,
            variance_s=s_series,
            prob_0=probs0,
            prob_1=probs1,
            predicted_class=-cls,
        )
        return self

    # ---- plotting: presentation only (out of scope of the hot path); thin optional pass-through ------------------
    # Without matplotlib (it is not part of this image) every plot_* method prints a note and returns, so demo drivers
    # in the style of graph_test.py run to the end; with it they draw the same quantities as the reference's methods.
    def _pyplot(self):
        try:
            import matplotlib.pyplot as pl
        except ImportError:
            print("> matplotlib not available: plot skipped")
            return None
        return pl

    @staticmethod
    def _save(pl, filename):
        if isinstance(filename, str):
            pl.savefig(filename, bbox_inches="tight")

    @staticmethod
    def int_to_list(nodes):
        """GPnet.py:566-570: a single node label becomes a one-element list."""
        return [nodes] if isinstance(nodes, (int, np.integer)) else nodes

    def plot_graph(self, filename=False):
        """Training / test / other nodes in red / green / blue on the cached layout (GPnet.py:378-433)."""
        pl = self._pyplot()
        if pl is None:
            return self
        pl.figure(figsize=[15, 9])
        pl.title("Graph")
        for nodes, color, label in ((self.training_nodes, "r", "training nodes"), (self.test_nodes, "g", "test nodes"),
                                    (self.other_nodes, "b", "other nodes")):
            nx.draw_networkx_nodes(self.Graph, self.plot_pos, nodelist=list(nodes), node_color=color, node_size=200, label=label)
        nx.draw_networkx_edges(self.Graph, self.plot_pos, alpha=0.2)
        names = self.orig_labels_invdict if getattr(self, "relabel_nodes", False) and hasattr(self, "orig_labels_invdict") else None
        nx.draw_networkx_labels(self.Graph, pos=self.plot_pos, labels=names, font_color="k")
        pl.legend()
        self._save(pl, filename)
        return self

    def _plot_draws(self, cov, mean, title, filename):
        pl = self._pyplot()
        if pl is None:
            return
        m = cov.shape[0]
        chol = np.linalg.cholesky(cov + 1e-6 * np.eye(m))
        draws = np.reshape(mean, (-1, 1)) + chol @ np.random.normal(size=(m, 5))
        pl.figure()
        pl.plot(self.test_nodes, draws)
        pl.title(title)
        pl.xlabel("nodes")
        pl.ylabel("values")
        self._save(pl, filename)

    def plot_prior(self, filename=False):
        """Five draws from the prior over the test nodes (GPnet.py:435-447)."""
        self._plot_draws(np.asarray(self.kstarstar), 0.0, "5 draws from the prior", filename)

    def plot_post(self, filename=False):
        """Five draws from the posterior over the test nodes (GPnet.py:449-467)."""
        v = np.asarray(self.v)               # L^-1 kstar^T
        self._plot_draws(np.asarray(self.kstarstar) - v.T @ v, np.asarray(self.fstar), "5 draws from the posterior", filename)

    def _plot_on_graph(self, title, train_colors, test_colors, filename):
        """Shared body of plot_predict_graph (GPnetRegressor.py:284-336, GPnetClassifier.py:280-346): training nodes as
        triangles, test nodes as squares, both coloured through self.cmap; other nodes grey."""
        pl = self._pyplot()
        if pl is None:
            return self
        pl.figure(figsize=[15, 9])
        pl.title(title)
        self.gen_cmap()
        common = dict(node_size=200, cmap=self.cmap, vmin=self.vmin, vmax=self.vmax)
        nx.draw_networkx_nodes(self.Graph, self.plot_pos, nodelist=list(self.training_nodes), node_color=train_colors, node_shape="v", **common)
        nx.draw_networkx_nodes(self.Graph, self.plot_pos, nodelist=list(self.other_nodes), node_color="gray", node_size=200)
        nx.draw_networkx_nodes(self.Graph, self.plot_pos, nodelist=list(self.test_nodes), node_color=test_colors, node_shape="s", **common)
        nx.draw_networkx_edges(self.Graph, self.plot_pos, alpha=0.2)
        sm = pl.cm.ScalarMappable(cmap=self.cmap, norm=pl.Normalize(vmin=self.vmin, vmax=self.vmax))
        sm.set_array([])
        pl.colorbar(sm, ax=pl.gca())
        self._save(pl, filename)
        return self
