"""Host-side engine shared by the drop-in classes: one library context per process per GPU, graphs
uploaded once and re-bound only when another model used the context in between (SURVEY A-17: the
reference rebuilds the Laplacian from networkx on every kernel() call, GPnet.py:329)."""
from __future__ import annotations

import os

import numpy as np

from ._lib import KERNEL_IDS, Context

_CONTEXTS: dict[int, Context] = {}


def default_device() -> int:
    """LOCAL_RANK under torchrun (one process per GPU), else GPNET_B200_DEVICE, else 0."""
    for key in ("GPNET_B200_DEVICE", "LOCAL_RANK"):
        if key in os.environ:
            return int(os.environ[key])
    return 0


def get_context(device: int | None = None) -> Context:
    device = default_device() if device is None else int(device)
    ctx = _CONTEXTS.get(device)
    if ctx is None:
        ctx = Context(device)          # raises if the .so is missing or no B200 is visible: no CPU fallback
        ctx._owner = None
        ctx._nodes_key = None
        ctx._kernel_key = None
        ctx._predict_token = None
        _CONTEXTS[device] = ctx
    return ctx


def register_context(ctx: Context) -> Context:
    """Adopt an externally created context (e.g. one bound to a torch stream) as THE context of its device."""
    for attr in ("_owner", "_nodes_key", "_kernel_key", "_predict_token"):
        if not hasattr(ctx, attr):
            setattr(ctx, attr, None)
    _CONTEXTS[ctx.device] = ctx
    return ctx


def graph_csr(G, weight="weight"):
    """CSR adjacency in ``list(G)`` order (what nx.normalized_laplacian_matrix(G) uses at GPnet.py:329);
    a self-loop contributes once to its row, undirected edges appear in both rows, parallel edges of a MultiGraph
    are summed into one entry (as ``nx.to_scipy_sparse_array`` does).  Directed graphs raise the same
    ``NetworkXNotImplemented`` the reference's ``nx.normalized_laplacian_matrix`` call raises for them."""
    if G.is_directed():
        import networkx as nx
        raise nx.NetworkXNotImplemented("not implemented for directed type")
    nodelist = list(G)
    pos = {u: i for i, u in enumerate(nodelist)}
    N = len(nodelist)
    src, dst, wts = [], [], []
    for u, v, d in G.edges(data=True):
        w = float(d.get(weight, 1.0))
        i, j = pos[u], pos[v]
        src.append(i); dst.append(j); wts.append(w)
        if i != j:
            src.append(j); dst.append(i); wts.append(w)
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    wts = np.asarray(wts, dtype=np.float64)
    order = np.lexsort((dst, src))
    src, dst, wts = src[order], dst[order], wts[order]
    if len(src) > 1:            # coalesce duplicate (i, j) entries: one CSR entry per matrix entry
        first = np.ones(len(src), dtype=bool)
        first[1:] = (src[1:] != src[:-1]) | (dst[1:] != dst[:-1])
        if not first.all():
            group = np.cumsum(first) - 1
            wts = np.bincount(group, weights=wts)
            src, dst = src[first], dst[first]
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(indptr, src + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr.astype(np.int32), dst.astype(np.int32), wts, nodelist, pos


class PredictToken:
    """Marks the device buffers a regressor's predict() left behind (k, kstar, L, v, kstarstar are fetched lazily)."""

    def __init__(self, model):
        import weakref
        self.model = weakref.ref(model)


def invalidate_predict(ctx):
    """Call before ANY device call that may reuse the buffers of the last predict(): the model that owns them copies its
    not-yet-fetched attributes to the host first, so they stay readable afterwards like the reference's host arrays
    (GPnetRegressor.py:89-128 keeps k, kstar, kstarstar, L, v, V as ndarray attributes forever)."""
    tok = ctx._predict_token
    ctx._predict_token = None
    if tok is not None:
        model = tok.model()
        if model is not None:
            ctx._predict_token = tok          # the fetches below check it
            try:
                model._materialize()
            finally:
                ctx._predict_token = None


def lane_sm_budget(num_sms: int, lanes: int) -> int:
    """SMs each of ``lanes`` concurrent contexts may fill with co-resident factorisation kernels (0: the whole chip).  An even split
    for two lanes; from three lanes on a mild over-subscription (40 % of the chip each: a cooperative launch that does not fit
    simply waits) measured best at C3: 3 lanes with 40 / 49 / 60 / 74 SMs: 738 / 804 / 820 / 817 evals/s; 2 lanes with 60 / 74 / 90:
    689 / 733 / 730; 4 lanes with 37 / 49 / 60: 702 / 762 / 784."""
    if lanes <= 1:
        return 0
    if lanes == 2:
        return num_sms // 2
    return max(num_sms // lanes, (2 * num_sms) // 5)


def evaluate_batch(ctxs, kind: int, thetas, t, want_grad=True):
    """(-logp, -grad) rows and Cholesky infos for a batch of independent log-parameter vectors, dealt round robin to the
    lane contexts ``ctxs`` (all bound to the same graph / node sets): every lane always has one evaluation in flight, so on
    the asynchronous route the evaluations overlap on the GPU.  ``t``: host targets, or None for the resident ones.
    Parameter-domain violations raise the reference's AssertionError after the lanes have drained."""
    thetas = [np.asarray(th, dtype=np.float64).reshape(-1) for th in thetas]
    B, J = len(thetas), max(len(ctxs), 1)
    P = len(thetas[0]) if B else 0
    out = np.zeros((B, 1 + P), dtype=np.float64)
    infos = np.zeros(B, dtype=np.int64)
    pending = [None] * J
    bad = 0
    for b in range(B + J):
        lane = b % J
        if pending[lane] is not None:
            o, info = ctxs[lane].gpr_logp_grad_end()
            out[pending[lane], : len(o)] = o
            infos[pending[lane]] = info
            pending[lane] = None
        if b < B and not bad:
            st = ctxs[lane].gpr_logp_grad_begin(kind, thetas[b], t, want_grad)
            if st in (-100, -101):
                bad = st
            else:
                pending[lane] = b
    raise_domain(bad)
    return out, infos


class GraphEngine:
    """Binds one model's graph + node sets to the shared context on demand."""

    def __init__(self, G, device: int | None = None):
        self.indptr, self.indices, self.weights, self.nodelist, self.pos = graph_csr(G)
        self.N = len(self.nodelist)
        self.device = device
        self._ctx = None
        self._pos_cache = {}
        self._lanes = []
        self._lanes_key = None
        self.spectral = False          # kernels as V r(lam) V^T from one eigendecomposition of L~ (GPnetBase.use_spectral)
        self.spectral_info = None

    def lanes(self, count: int, train_pos, test_pos=()):
        """``count`` contexts bound to this graph and these node sets for concurrent evaluations: lane 0 is the shared
        context, the others are private to this engine (their buffers are allocated on first use).  Each lane may fill
        1/count of the SMs with co-resident factorisation kernels."""
        ctx = self.bind(train_pos, test_pos)
        count = max(1, int(count))
        key = (tuple(train_pos), tuple(test_pos or ()))
        while len(self._lanes) < count - 1:
            self._lanes.append(Context(ctx.device))
            self._lanes_key = None
        extra = self._lanes[: count - 1]
        if self._lanes_key != key:
            for lane in self._lanes:
                lane.set_graph(self.indptr, self.indices, self.weights)
                lane.set_nodes(train_pos, test_pos or ())
            self._lanes_key = key
        share = lane_sm_budget(ctx.num_sms(), count)
        for lane in [ctx] + extra:
            lane.set_sm_budget(share)
        return [ctx] + extra

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = get_context(self.device)
        return self._ctx

    def bind(self, train_pos=None, test_pos=None) -> Context:
        ctx = self.ctx
        if ctx._owner is not self:
            invalidate_predict(ctx)
            ctx.set_graph(self.indptr, self.indices, self.weights)
            ctx._owner = self
            ctx._nodes_key = None
            ctx._kernel_key = None
            if self.spectral:
                self.spectral_info = ctx.spectral_prepare()       # once per bound graph (again if another model took the context)
        if bool(getattr(ctx, "_spectral_on", False)) != self.spectral:
            if self.spectral:
                self.spectral_info = ctx.spectral_prepare()
            ctx.spectral_enable(self.spectral)
            ctx._spectral_on = self.spectral
            ctx._kernel_key = None
        if train_pos is not None:
            key = (tuple(train_pos), tuple(test_pos or ()))
            if ctx._nodes_key != key:
                invalidate_predict(ctx)       # a pending predict() snapshot needs the node sets it was computed with
                ctx.set_nodes(train_pos, test_pos or ())
                ctx._nodes_key = key
        return ctx

    def positions(self, nodes, relabelled: bool):
        """graph positions of ``nodes``, de-duplicated and ascending (GPnet.py:305-323 set algebra, A-5).  The optimiser and
        landscape loops pass the same node list thousands of times: the mapping is memoised on the tuple of nodes."""
        try:
            key = (tuple(nodes), bool(relabelled))
            hit = self._pos_cache.get(key)
        except TypeError:            # unhashable node labels
            key, hit = None, None
        if hit is not None:
            return hit
        out = sorted({int(x) for x in nodes}) if relabelled else sorted({self.pos[x] for x in nodes})
        if out and (out[0] < 0 or out[-1] >= self.N):      # the reference's np.delete / fancy indexing raises IndexError
            bad = out[0] if out[0] < 0 else out[-1]
            raise IndexError(f"index {bad} is out of bounds for axis 0 with size {self.N}")
        if key is not None:
            if len(self._pos_cache) > 8:
                self._pos_cache.clear()
            self._pos_cache[key] = out
        return out

    def build_kernel(self, kerneltype: str, theta, want_grad=False):
        """Full N x N kernel resident on the device for these log-parameters (cached per theta)."""
        ctx = self.bind()
        th = np.asarray(theta, dtype=np.float64).reshape(-1)
        key = (kerneltype, th.tobytes(), bool(want_grad))
        if ctx._kernel_key == key:
            return ctx
        ctx._kernel_key = None
        invalidate_predict(ctx)
        st = ctx.kernel_build(KERNEL_IDS[kerneltype], th, want_grad)
        raise_domain(st)
        ctx._kernel_key = key
        return ctx

    def invalidate_kernel(self):
        if self._ctx is not None:
            self._ctx._kernel_key = None


def raise_domain(st: int):
    """Parameter-domain statuses of the C-ABI -> the reference's AssertionErrors (GPnet.py:339,344)."""
    if st == -100:
        raise AssertionError("Lambda must be < 1")
    if st == -101:
        raise AssertionError("a must be >=2")
