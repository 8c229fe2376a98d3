"""CPU tier: host-side logic of the drop-in (graph -> CSR, node bookkeeping, sharding, the C-ABI surface)."""
import os
import re
import subprocess
import sys

import networkx as nx
import numpy as np
import pandas as pd
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    """The .so loads without a GPU and exports every function include/gpnet_b200.h declares."""
    from gpnet_b200 import _lib
    lib = _lib.load_library()
    header = open(os.path.join(ROOT, "include", "gpnet_b200.h")).read()
    declared = set(re.findall(r"\b(gpn_[a-z0-9_]+)\s*\(", header))
    declared -= {"gpn_ctx"}
    assert declared, "no declarations found"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gpnet_b200._lib import Context, GpnetError
    with pytest.raises(GpnetError):
        Context(0)


def test_graph_csr_matches_networkx_adjacency():
    from gpnet_b200.engine import graph_csr
    G = nx.barabasi_albert_graph(50, 3, seed=3)
    G.add_edge(4, 4)
    G[0][1]["weight"] = 0.25
    indptr, indices, w, nodelist, pos = graph_csr(G)
    import scipy.sparse as sp
    A = sp.csr_array((w, indices, indptr), shape=(50, 50)).toarray()
    assert np.array_equal(A, nx.to_numpy_array(G, nodelist=nodelist))
    assert nodelist == list(G)


def test_graph_csr_sums_parallel_edges_and_rejects_directed_graphs():
    """One CSR entry per matrix entry: a MultiGraph's parallel edges are summed like nx.to_scipy_sparse_array does (the
    Laplacian scatter kernel writes one value per entry); directed graphs raise what the reference's
    nx.normalized_laplacian_matrix call raises for them (GPnet.py:329)."""
    from gpnet_b200.engine import graph_csr
    import scipy.sparse as sp
    G = nx.MultiGraph()
    G.add_nodes_from(range(5))
    G.add_edge(0, 1, weight=0.5)
    G.add_edge(0, 1, weight=2.0)
    G.add_edge(1, 0)                       # default weight 1
    G.add_edge(2, 3)
    G.add_edge(3, 3, weight=1.5)
    G.add_edge(3, 3, weight=0.25)
    indptr, indices, w, nodelist, _ = graph_csr(G)
    A = sp.csr_array((w, indices, indptr), shape=(5, 5))
    assert A.nnz == len(indices) == 5       # (0,1) (1,0) (2,3) (3,2) (3,3)
    assert np.array_equal(A.toarray(), nx.to_numpy_array(G, nodelist=nodelist))
    assert np.allclose(nx.normalized_laplacian_matrix(G).toarray()[0, 1], -3.5 / np.sqrt(3.5 * 3.5))
    with pytest.raises(nx.NetworkXNotImplemented):
        graph_csr(nx.DiGraph([(0, 1), (1, 2)]))


def test_positions_outside_the_graph_raise_index_error():
    """relabelled graphs take int(label) as the position; labels outside [0, N) must not reach the device gather."""
    m = _model()
    with pytest.raises(IndexError):
        m._positions([0, 1, 30])
    with pytest.raises(IndexError):
        m._positions([-1, 2])
    assert m._positions([29, 0]) == [0, 29]


def test_product_does_not_import_oracle():
    for fn in os.listdir(os.path.join(ROOT, "gpnet_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "gpnet_b200", fn)).read()
            assert "oracle" not in src, fn


def _model(**kw):
    import gpnet_b200
    G = nx.grid_2d_graph(6, 5)
    return gpnet_b200.GPnetRegressor(Graph=G, ntrain=10, ntest=8, theta=np.log([0.6, 0.01]), relabel_nodes=True, **kw)


def test_constructor_bookkeeping(capsys):
    m = _model(seed=3)
    assert len(m.training_nodes) == 10 and len(m.test_nodes) == 8 and len(m.other_nodes) == 12
    assert m.training_nodes == sorted(m.training_nodes)
    assert not (set(m.training_nodes) & set(m.test_nodes))
    assert set(m.training_nodes) | set(m.test_nodes) | set(m.other_nodes) == set(range(30))
    assert m.orig_labels_dict[(1, 0)] == 5 and m.orig_labels_invdict[5] == (1, 0)
    assert m.pvtdist[29] == 9 and m.pivot_flag
    assert list(m.training_values.index) == m.training_nodes
    m2 = _model(seed=3)
    assert m2.training_nodes == m.training_nodes          # seeded
    assert "Assigning Nodes Randomly" in capsys.readouterr().out


def test_constructor_errors_and_series_values():
    import gpnet_b200
    G = nx.grid_2d_graph(3, 3)
    with pytest.raises(ValueError):
        gpnet_b200.GPnetRegressor(Graph=G, ntrain=6, ntest=6, relabel_nodes=True)
    tv = pd.Series([1.0, 2.0, 3.0], index=[0, 4, 8])
    m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=[0, 4, 8], test_nodes=[1, 2], training_values=tv, relabel_nodes=True)
    assert m.training_values is tv and not m.pivot_flag
    y = pd.Series([1, -1, 1], index=[0, 4, 8])
    c = gpnet_b200.GPnetClassifier(Graph=G, training_nodes=[0, 4, 8], test_nodes=[1, 2], training_values=y, relabel_nodes=True)
    assert c.training_values is y
    c2 = gpnet_b200.GPnetClassifier(Graph=G, training_nodes=[0, 4, 8], test_nodes=[1, 2], relabel_nodes=True)
    assert set(np.unique(c2.training_values.values)) <= {-1, 1} and c2.pivot_flag


def test_positions_follow_graph_order_not_caller_order():
    """A-5 / GPnet.py:305-323: un-relabelled graphs map node names through orig_labels_dict; order is ascending."""
    import gpnet_b200
    G = nx.grid_2d_graph(3, 3)
    m = gpnet_b200.GPnetRegressor(Graph=G, training_nodes=[(2, 2), (0, 1), (1, 0)], test_nodes=[(0, 0)], relabel_nodes=False)
    assert m._positions([(2, 2), (0, 1), (1, 0), (0, 1)]) == [1, 3, 8]


def test_dropin_module_aliases():
    import gpnet_b200
    gpnet_b200.install_dropin()
    from GPnetRegressor import GPnetRegressor
    from GPnetClassifier import GPnetClassifier, LAMBDAS, COEFS
    assert GPnetRegressor is gpnet_b200.GPnetRegressor and GPnetClassifier is gpnet_b200.GPnetClassifier
    assert LAMBDAS.shape == (5,) and COEFS.shape == (5, 1)


def test_kernel_type_assert_is_raised_on_host():
    m = _model()
    m.kerneltype = "matern"
    with pytest.raises(AssertionError, match="kerneltype not implemented"):
        m.kernel([0], [1], np.log([0.5, 0.1]))


class _FakeCtx:
    """Stands in for the device context so the host-side mapping of LAPACK-style info codes can be tested on CPU."""
    _owner = None; _nodes_key = None; _kernel_key = None; _predict_token = None; device = 0

    def __init__(self, info_k=0, info_kss=0):
        self.info_k, self.info_kss = info_k, info_kss

    def set_graph(self, *a): pass
    def set_nodes(self, tr, te=()): self.n_train, self.n_test = len(tr), len(te)

    def gpr_logp_grad(self, kind, theta, t, want_grad):
        return 0, np.array([12.5] + [0.25] * len(theta)), self.info_k

    def gpr_predict(self, kind, theta, t):
        n, m = self.n_train, self.n_test
        self.generation = getattr(self, "generation", 0) + 1
        return 0, dict(alpha=np.zeros(n), fstar=np.ones(m), s=np.ones(m), kstarstar_diag=np.ones(m), info_k=self.info_k,
                       info_kss=self.info_kss)

    def gpr_fetch(self, which):
        n, m = self.n_train, self.n_test
        shape = {"k": (n, n), "kstar": (m, n), "L": (n, n), "v": (n, m), "kstarstar": (m, m)}[which]
        self.fetched = getattr(self, "fetched", []) + [which]
        return np.full(shape, float(self.generation))

    def kernel_build(self, kind, theta, want_grad=False):
        return 0


def test_cholesky_failure_maps_to_reference_behaviour(capsys):
    """A-8: logPosterior returns -inf when the Cholesky fails (GPnetRegressor.py:140-143); A-7: the PD gates of predict
    print, set the flag and return self without training (GPnetRegressor.py:112-121)."""
    m = _model()
    m._engine._ctx = _FakeCtx(info_k=3)
    assert m.logPosterior(m.theta, m.training_nodes, m.training_values) == -np.inf
    assert m.predict() is m and m.k_not_posdef_flag and not m.kstar_not_posdef_flag and not m.is_trained
    assert "K not positive definite, aborting..." in capsys.readouterr().out
    m._engine._ctx = _FakeCtx(info_kss=2)
    assert m.logPosterior(m.theta, m.training_nodes, m.training_values) == 12.5
    assert np.array_equal(m.gradLogPosterior(m.theta, m.training_nodes, m.training_values), [0.25, 0.25])
    assert m.predict() is m and m.kstar_not_posdef_flag and not m.is_trained
    assert "K** not positive definite, aborting..." in capsys.readouterr().out
    m._engine._ctx = _FakeCtx()
    assert m.predict().is_trained and list(m.df["fstar"][m.test_nodes]) == [1.0] * 8


def test_predict_attributes_survive_later_device_calls():
    """k, kstar, kstarstar, L, v, V are fetched lazily, but stay readable after logp() / kernel() / another model's
    predict() reused the device buffers -- their owner snapshots them right before (the reference keeps them as host arrays,
    GPnetRegressor.py:89-128, so gpr.predict(); gpr.logp(); gpr.plot_post() works there)."""
    m = _model()
    ctx = _FakeCtx()
    m._engine._ctx = ctx
    m.predict()
    assert not getattr(ctx, "fetched", [])                      # nothing copied yet: lazy
    assert m.k.shape == (10, 10) and ctx.fetched == ["k"]       # first read fetches one array
    m.logp()                                                    # reuses the buffers: the rest is snapshotted first
    assert sorted(ctx.fetched) == ["L", "k", "kstar", "kstarstar", "v"]
    assert m.L.shape == (10, 10) and m.v.shape == (10, 8) and m.kstarstar.shape == (8, 8) and m.V.shape == (8, 8)
    assert np.all(m.v == 1.0)
    other = _model(seed=5)                                      # a second model on the same context
    other._engine._ctx = ctx
    m.predict()                                                 # generation 2
    other.predict()                                             # must snapshot m's generation-2 arrays before running
    assert np.all(m.kstar == 2.0) and np.all(other.kstar == 3.0)


def test_workspace_bytes_matches_the_measured_footprints():
    """gpn_workspace_bytes is host arithmetic over the buffers a context can allocate (an upper bound: regressor and
    classifier scratch together): C5 (N = 32768, n = m = 16384, diffusion) used 106 GB of HBM on the B200
    (tools/c5_check.py), the headline configuration a few GB."""
    from gpnet_b200._lib import workspace_bytes
    c5 = workspace_bytes(32768, 16384, 16384, "diffusion")
    assert 1.06e11 <= c5 < 1.25e11
    c3 = workspace_bytes(8192, 4096, 4096, "regularized_laplacian")
    assert 2.5e9 < c3 < 6e9
    assert workspace_bytes(8192, 4096, 4096, "diffusion") > workspace_bytes(8192, 4096, 4096, "pstep_walk") > c3
    assert workspace_bytes(0, 1, 1, "diffusion") == -1


def test_cyclic_shards_cover_the_grid():
    from gpnet_b200.sharded import cyclic_slice
    total = 64 * 64
    for world in (1, 2, 4, 8, 3):
        seen = np.zeros(total, dtype=int)
        for r in range(world):
            b, e, s = cyclic_slice(total, r, world)
            seen[b:e:s] += 1
        assert np.all(seen == 1)


def test_kernel_group_shards_cover_the_grid_and_keep_groups_together():
    """Grid points that differ only in the noise entry share a kernel: whole groups go to one rank, every index to exactly
    one rank; without sharing (p-step with P = 2, noise not on the grid) the split is the plain cyclic one."""
    from gpnet_b200.sharded import kernel_group_keys, shard_indices
    n1, n2 = 7, 5
    for axidx, P, kt, ngroups in (([0, 1], 2, "diffusion", n1), ([1, 0], 2, "regularized_laplacian", n2), ([1, 1], 2, "diffusion", 1),
                                  ([0, 1], 2, "pstep_walk", n1 * n2), ([0, 1], 3, "pstep_walk", n1 * n2), ([0, 2], 3, "pstep_walk", n1)):
        keys = kernel_group_keys(n1, n2, axidx, P, kt)
        assert len(np.unique(keys)) == ngroups
        for world in (1, 2, 3, 8):
            seen = np.zeros(n1 * n2, dtype=int)
            owner = {}
            for r in range(world):
                idx = shard_indices(keys, r, world)
                seen[idx] += 1
                for k in np.unique(keys[idx]):
                    assert owner.setdefault(int(k), r) == r
            assert np.all(seen == 1)
    keys = kernel_group_keys(4, 4, [0, 1], 2, "pstep_walk")
    assert np.array_equal(shard_indices(keys, 1, 3), np.arange(1, 16, 3))


def test_cost_balanced_sharding_of_the_c4_grid():
    """C4 (64 x 64 diffusion grid): kernel groups (one per beta) cost more the higher their Pade order; dealt longest first
    to the least loaded rank every rank gets the same estimated cost (within one group), whole groups stay together."""
    from gpnet_b200.sharded import kernel_group_costs, kernel_group_keys, pade_cost, shard_indices
    ax1 = np.linspace(np.log(0.01), np.log(0.99), 64)
    ax2 = np.linspace(np.log(0.001), np.log(10.0), 64)
    keys = kernel_group_keys(64, 64, [0, 1], 2, "diffusion")
    costs = kernel_group_costs(keys, 64, 64, [0, 1], 2, "diffusion", ax1, ax2)
    assert len(costs) == 64 and costs[0] < costs[63] and pade_cost(0.001) < pade_cost(0.5) < pade_cost(3.0)
    for world in (2, 4, 8, 3):
        shards = [shard_indices(keys, r, world, costs) for r in range(world)]
        seen = np.zeros(64 * 64, dtype=int)
        loads = []
        for sh in shards:
            seen[sh] += 1
            groups = np.unique(keys[sh])
            assert all(np.sum(keys[sh] == g) == 64 for g in groups)          # whole groups
            loads.append(sum(costs[int(g)] for g in groups))
        assert np.all(seen == 1)
        assert max(loads) - min(loads) <= max(costs.values()) + 1e-9
    # the regularized Laplacian costs the same everywhere: uniform costs
    assert set(kernel_group_costs(keys, 64, 64, [0, 1], 2, "regularized_laplacian", ax1, ax2).values()) == {1.0}


WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
from gpnet_b200.sharded import allgather_rows, cyclic_slice, sharded_multistart
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT={port!r}, RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group("gloo", rank=rank, world_size=world)
total, width = 37, 3
b, e, s = cyclic_slice(total, rank, world)
idx = np.arange(b, e, s)
local = np.stack([idx * 1.0, idx * 2.0 + 0.5, -idx * 1.0], axis=1)
full = allgather_rows(local, total, rank, world, dist)
exp = np.stack([np.arange(total) * 1.0, np.arange(total) * 2.0 + 0.5, -np.arange(total) * 1.0], axis=1)
assert np.array_equal(full, exp), (rank, full[:5])
from gpnet_b200.sharded import kernel_group_keys, shard_indices      # ragged shards: 5 kernel groups of 4 points on 2 ranks
keys = kernel_group_keys(5, 4, [0, 1], 2, "diffusion")
shards = [shard_indices(keys, r, world) for r in range(world)]
mine = shards[rank]
full = allgather_rows(np.stack([mine * 1.0, mine * 3.0], axis=1), 20, rank, world, dist, None, shards)
assert np.array_equal(full, np.stack([np.arange(20) * 1.0, np.arange(20) * 3.0], axis=1)), (rank, full)

class Fake:                                    # stands in for a model: value and gradient are functions of theta
    training_nodes = [0]; training_values = [0.0]
    def logPosterior_and_grad(self, th, *a):
        return float(np.sum(th ** 2)), 2 * np.asarray(th)
thetas = np.arange(14, dtype=float).reshape(7, 2)
out = sharded_multistart(Fake(), thetas)
assert np.allclose(out[:, 0], (thetas ** 2).sum(1)) and np.allclose(out[:, 1:], 2 * thetas)
# a parameter-domain failure on ONE rank reaches every rank as the same exception (no rank is left waiting in the all-gather)
from gpnet_b200.sharded import agree_on_error
try:
    agree_on_error(AssertionError("Lambda must be < 1") if rank == 1 else None, world, dist)
    raise SystemExit("agree_on_error did not raise on rank %d" % rank)
except AssertionError as e:
    assert "Lambda must be < 1" in str(e)
agree_on_error(None, world, dist)             # no failure anywhere: no exception
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_allgather_world_size_2_gloo(tmp_path):
    """N > 1 path on CPU: two gloo ranks reassemble cyclic shards into the identical full array."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=str(29500 + os.getpid() % 2000)))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()


def test_lane_sm_budgets():
    """SM budget per lane (engine.lane_sm_budget): whole chip for one lane, an even split for two, a mild over-subscription
    (40 % each) from three lanes on -- never less than the even share."""
    from gpnet_b200.engine import lane_sm_budget
    assert lane_sm_budget(148, 1) == 0 and lane_sm_budget(148, 0) == 0
    assert lane_sm_budget(148, 2) == 74
    assert lane_sm_budget(148, 3) == 59 and lane_sm_budget(148, 4) == 59
    for lanes in range(2, 9):
        assert lane_sm_budget(148, lanes) >= 148 // lanes
