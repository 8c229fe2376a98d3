"""Canonical synthetic inputs, shared by the golden generator, the tests and
bench.py's CPU legs.  TEST INFRASTRUCTURE (see oracle/__init__.py); pure host code, no reference import."""
import math

import networkx as nx
import numpy as np


def relabel(G):
    return nx.relabel_nodes(G, dict(zip(G, range(len(G)))))


def split(N, n, m, seed=0):
    perm = np.random.RandomState(seed).permutation(N)
    return sorted(int(x) for x in perm[:n]), sorted(int(x) for x in perm[n:n + m])


def bfs_dist(G, src=0):
    d = nx.single_source_shortest_path_length(G, src)
    return np.array([d[i] for i in range(len(G))], dtype=np.float64)


def grid_case(a, b, n, m, freq=0.5):
    """C1/C5 family: 2-D grid, sine field of the BFS distance from node 0 (graph_test.py:40 idiom)."""
    G = relabel(nx.grid_2d_graph(a, b))
    train, test = split(len(G), n, m)
    t = np.sin(freq * bfs_dist(G))
    return G, train, test, t


def ba_case(N, n, m, mattach=3, freq=1.5):
    """C2 family: Barabasi-Albert, +-1 labels."""
    G = nx.barabasi_albert_graph(N, mattach, seed=0)
    train, test = split(N, n, m)
    y = np.where(np.sin(freq * bfs_dist(G)) > 0, 1, -1).astype(np.int64)
    return G, train, test, y


def rgg_case(N, n, m, mean_deg=20.0):
    """C3/C4 family: random geometric graph, smooth field on node positions."""
    r = math.sqrt(mean_deg / (math.pi * N))
    G = nx.random_geometric_graph(N, r, seed=0)
    train, test = split(N, n, m)
    pos = np.array([G.nodes[i]["pos"] for i in range(N)])
    t = np.sin(4 * np.pi * pos[:, 0]) * np.cos(4 * np.pi * pos[:, 1])
    return G, train, test, t


def rgg_csr_fast(N, mean_deg=20.0, seed=0):
    """Same family as rgg_case but built with a KD-tree straight into CSR (no networkx object):
    used where N is large (bench / full-size property tests).  NOT the same node set as networkx's
    generator -- only for size-independent properties and throughput runs."""
    from scipy.spatial import cKDTree
    rng = np.random.RandomState(seed)
    pos = rng.rand(N, 2)
    r = math.sqrt(mean_deg / (math.pi * N))
    pairs = cKDTree(pos).query_pairs(r, output_type="ndarray")
    i = np.concatenate([pairs[:, 0], pairs[:, 1]])
    j = np.concatenate([pairs[:, 1], pairs[:, 0]])
    order = np.lexsort((j, i))
    i, j = i[order], j[order]
    indptr = np.zeros(N + 1, dtype=np.int32)
    np.add.at(indptr, i + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    t = np.sin(4 * np.pi * pos[:, 0]) * np.cos(4 * np.pi * pos[:, 1])
    return indptr, j.astype(np.int32), np.ones(len(j)), pos, t


def grid_csr_fast(a, b):
    """CSR of nx.grid_2d_graph(a, b) relabelled row-major, without building the networkx object."""
    N = a * b
    idx = np.arange(N).reshape(a, b)
    e = [(idx[:-1, :].ravel(), idx[1:, :].ravel()), (idx[:, :-1].ravel(), idx[:, 1:].ravel())]
    i = np.concatenate([e[0][0], e[0][1], e[1][0], e[1][1]])
    j = np.concatenate([e[0][1], e[0][0], e[1][1], e[1][0]])
    order = np.lexsort((j, i))
    i, j = i[order], j[order]
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(indptr, i + 1, 1)
    return np.cumsum(indptr).astype(np.int32), j.astype(np.int32), np.ones(len(j))
