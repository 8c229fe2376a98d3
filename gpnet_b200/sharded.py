"""Sharding of independent hyper-parameter evaluations (the only naturally parallel axis of the path,
SURVEY.md 8e): landscape grid points (GPnet.py:539-552) and multi-start evaluations are split cyclically
over the ranks of a torch.distributed job (one process per GPU) and the (1+P) doubles per point are
all-gathered (NCCL over NVLink on GPUs; gloo in the CPU tests).  A single factorisation never leaves its GPU.
"""
from __future__ import annotations

import numpy as np


def dist_info(group=None):
    """(rank, world_size, dist module or None)."""
    try:
        import torch.distributed as dist
    except ImportError:   # pragma: no cover
        return 0, 1, None
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group), dist
    return 0, 1, None


def cyclic_slice(total: int, rank: int, world: int):
    """Flat indices handled by ``rank`` when no two grid points share a kernel: rank, rank+world, ...  (cost depends on
    beta through the Pade order only, so striding gives every rank the same mix)."""
    return rank, total, world


def kernel_group_keys(n1: int, n2: int, axidx, P: int, kerneltype: str) -> np.ndarray:
    """Key of the kernel every flat grid index (i1*n2 + i2) needs: points that differ only in the noise entry
    ``theta[P-1]`` (added to every kernel entry, quirk A-3) share it.  Mirrors ``gpn_landscape_points`` (csrc/gp.cu)."""
    flat = np.arange(n1 * n2, dtype=np.int64)
    last = P - 1
    share = P >= 2 and not (kerneltype == "pstep_walk" and P == 2) and (axidx[0] == last or axidx[1] == last)
    if not share:
        return flat
    if axidx[1] == last:
        return np.zeros_like(flat) if axidx[0] == last else flat // n2
    return flat % n2


def pade_cost(beta: float) -> float:
    """Relative cost of one diffusion kernel build as a function of beta = exp(theta0): N^3 products of the Pade order the
    scaling-and-squaring expm picks (Al-Mohy & Higham thresholds on ||beta L~|| <~ 2 beta) plus the SPD solve Q X = P
    ((2 (pi_m + s) + 7/3) N^3).  A proxy for load balancing only."""
    eta = 2.0 * beta
    if eta < 1.495585217958292e-2:
        prods = 1
    elif eta < 2.539398330063230e-1:
        prods = 2
    elif eta < 9.504178996162932e-1:
        prods = 3
    elif eta < 2.097847961257068:
        prods = 4
    else:
        prods = 6 + max(0, int(np.ceil(np.log2(eta / 4.25))))
    return 2.0 * (prods + 1) + 7.0 / 3.0


def kernel_group_costs(keys: np.ndarray, n1: int, n2: int, axidx, P: int, kerneltype: str, ax1, ax2, theta=None) -> dict:
    """Estimated cost of every kernel group (key -> cost): the diffusion kernel's Pade order depends on beta, so
    groups are not equally expensive; the other kernels cost the same for every theta."""
    uniq, first = np.unique(keys, return_index=True)
    costs = {}
    for k, f in zip(uniq, first):
        c = 1.0
        if kerneltype == "diffusion":
            i1, i2 = divmod(int(f), n2)
            th0 = None
            if axidx[1] == 0:
                th0 = ax2[i2]
            elif axidx[0] == 0:
                th0 = ax1[i1]
            elif theta is not None:
                th0 = np.asarray(theta, dtype=np.float64).reshape(-1)[0]
            if th0 is not None:
                c = pade_cost(float(np.exp(th0)))
        costs[int(k)] = c
    return costs


def shard_indices(keys: np.ndarray, rank: int, world: int, costs: dict | None = None) -> np.ndarray:
    """Flat indices of ``rank``.  Whole kernel groups go to one rank, so that a kernel is built on exactly one GPU.  Without
    ``costs`` groups are dealt cyclically (group g -> rank g % world; with one point per group: the plain cyclic split).
    With ``costs`` (key -> estimated cost) they are dealt longest first to the least loaded rank (LPT), which balances
    grids whose kernels differ in Pade order."""
    uniq, inverse = np.unique(keys, return_inverse=True)
    if costs is None or world == 1:
        return np.nonzero(inverse % world == rank)[0].astype(np.int64)
    order = sorted(range(len(uniq)), key=lambda g: (-costs.get(int(uniq[g]), 1.0), g))
    load = [0.0] * world
    owner = np.zeros(len(uniq), dtype=np.int64)
    for g in order:
        r = min(range(world), key=lambda x: (load[x], x))
        owner[g] = r
        load[r] += costs.get(int(uniq[g]), 1.0)
    return np.nonzero(owner[inverse] == rank)[0].astype(np.int64)


def _comm_device(dist, group, device_index=None):
    """Device of the tensors a collective of this group needs: the GPU of the library context for NCCL, the host for gloo."""
    import torch
    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device() if device_index is None else int(device_index))
    return torch.device("cpu")


_ERR_CODES = {1: (AssertionError, "Lambda must be < 1"), 2: (AssertionError, "a must be >=2"),
              3: (RuntimeError, "hyper-parameter evaluation failed on another rank")}


def agree_on_error(err, world, dist, group=None, device_index=None):
    """Every rank learns whether ANY rank failed (max-reduce of a small code) and raises the same kind of exception, instead of
    the healthy ranks blocking forever in the all-gather that follows.  ``err``: the local exception or None."""
    code = 0
    if err is not None:
        code = 3
        if isinstance(err, AssertionError):
            code = 2 if str(err).startswith("a must be") else 1
    if world > 1 and dist is not None:
        import torch
        flag = torch.tensor([code], dtype=torch.int32, device=_comm_device(dist, group, device_index))
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        code = int(flag.item())
    if err is not None:
        raise err
    if code:
        exc, msg = _ERR_CODES[code]
        raise exc(msg)


def allgather_rows(local: np.ndarray, total: int, rank: int, world: int, dist, group=None, shards=None, device_index=None) -> np.ndarray:
    """Reassemble ``total`` rows from per-rank shards: ``shards[r]`` lists the row indices rank r holds (default: the cyclic
    split r, r+world, ...).  One all-gather of equal-sized (padded) blocks."""
    width = local.shape[1]
    if world == 1 or dist is None:
        return local
    import torch
    if shards is None:
        shards = [np.arange(r, total, world) for r in range(world)]
    per = max(max(len(sh) for sh in shards), 1)
    device = _comm_device(dist, group, device_index)
    send = torch.zeros((per, width), dtype=torch.float64, device=device)
    send[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    recv = torch.empty((world * per, width), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, per, width)
    out = np.empty((total, width), dtype=np.float64)
    for r in range(world):
        out[shards[r]] = recv[r, : len(shards[r])]
    return out


def _bound_context(model):
    from .engine import invalidate_predict
    ctx = model._engine.bind(model._positions(model.training_nodes), model._positions(model.test_nodes))
    ctx._kernel_key = None
    invalidate_predict(ctx)
    return ctx


def _targets(model):
    t = model.training_values
    return np.asarray(getattr(t, "values", t), dtype=np.float64).reshape(-1)


def evaluate_slice(model, theta, axidx, ax1, ax2, begin, end, stride, want_grad=False):
    """(count, 1+P) array of (lml, dlml/dtheta) for the flat grid indices begin, begin+stride, ... < end."""
    from ._lib import KERNEL_IDS
    from .engine import raise_domain
    th = model._theta_vec(theta).copy()
    classifier = not hasattr(model, "logPosterior_and_grad")
    st, out = _bound_context(model).landscape(KERNEL_IDS[model.kerneltype], classifier, th, axidx, ax1, ax2, _targets(model), begin, end,
                                              stride, want_grad)
    raise_domain(st)
    return out


def evaluate_points(model, theta, axidx, ax1, ax2, indices, want_grad=False):
    """(len(indices), 1+P) array of (lml, dlml/dtheta) for an explicit list of flat grid indices."""
    from ._lib import KERNEL_IDS
    from .engine import raise_domain
    th = model._theta_vec(theta).copy()
    classifier = not hasattr(model, "logPosterior_and_grad")
    from ._lib import Context
    n_train = len(model._positions(model.training_nodes))
    if classifier and not want_grad and getattr(model._engine, "spectral", False) and 0 < n_train <= Context.SPECTRAL_BATCH_MAX_N and len(indices):
        # the classifier's landscape: the Laplace Newton loop of every grid point of this rank in ONE launch, one CTA per point
        n2 = len(ax2)
        thetas = np.tile(th, (len(indices), 1))
        for row, flat in enumerate(np.asarray(indices, dtype=np.int64)):
            thetas[row, axidx[0]] = ax1[int(flat) // n2]
            thetas[row, axidx[1]] = ax2[int(flat) % n2]
        st, logq, iters, infos = _bound_context(model).gpc_newton_batch(KERNEL_IDS[model.kerneltype], thetas, _targets(model))
        raise_domain(st)
        out = np.zeros((len(indices), 1 + len(th)))
        out[:, 0] = logq
        out[infos != 0, 0] = np.nan          # np.linalg.cholesky raised inside the loop (gpn_landscape_points reports NaN too)
        return out
    if not classifier and getattr(model._engine, "spectral", False) and 0 < n_train <= Context.SPECTRAL_BATCH_MAX_N and len(indices):
        # spectral path, small training set: all grid points of this rank in ONE launch, one CTA per point
        n2 = len(ax2)
        thetas = np.tile(th, (len(indices), 1))
        for row, flat in enumerate(np.asarray(indices, dtype=np.int64)):
            thetas[row, axidx[0]] = ax1[int(flat) // n2]
            thetas[row, axidx[1]] = ax2[int(flat) % n2]      # the second axis wins when both name the same entry, like the loop of GPnet.py:545-549
        st, rows, infos = _bound_context(model).gpr_logp_grad_batch(KERNEL_IDS[model.kerneltype], thetas, _targets(model), want_grad)
        raise_domain(st)
        out = -rows
        if not want_grad:
            out[:, 1:] = 0.0
        out[infos > 0, 0] = np.inf          # logPosterior returned -inf (A-8) => lml = +inf
        return out
    st, out = _bound_context(model).landscape_points(KERNEL_IDS[model.kerneltype], classifier, th, axidx, ax1, ax2, _targets(model), indices,
                                                     want_grad)
    raise_domain(st)
    return out


def sharded_landscape(model, theta, axidx, ax1, ax2, want_grad=False, group=None) -> np.ndarray:
    """Full ``(len(ax1), len(ax2), 1+P)`` array on every rank: [..., 0] is the lml of GPnet.py:549,
    [..., 1:] its gradient when ``want_grad``.  Kernel groups (grid points that differ only in the noise entry) stay
    together on one rank; one all-gather of (1+P) doubles per point reassembles the grid."""
    rank, world, dist = dist_info(group)
    n1, n2 = len(ax1), len(ax2)
    total = n1 * n2
    P = len(model._theta_vec(theta))
    keys = kernel_group_keys(n1, n2, axidx, P, model.kerneltype)
    costs = kernel_group_costs(keys, n1, n2, axidx, P, model.kerneltype, ax1, ax2, theta) if len(np.unique(keys)) < total else None
    shards = [shard_indices(keys, r, world, costs) for r in range(world)]
    err, local = None, np.zeros((len(shards[rank]), 1 + P))
    try:
        local = evaluate_points(model, theta, axidx, ax1, ax2, shards[rank], want_grad)
    except Exception as e:      # e.g. a diffusion axis reaching exp(theta0) >= 1 on this rank's points only
        err = e
    dev = getattr(model._engine, "_ctx", None)
    dev = dev.device if dev is not None else None
    agree_on_error(err, world, dist, group, dev)
    full = allgather_rows(local, total, rank, world, dist, group, shards, dev)
    return full.reshape(n1, n2, -1)


def sharded_multistart(model, thetas, want_grad=True, group=None) -> np.ndarray:
    """(-logp, -grad) for every row of ``thetas`` (multi-start / finite-difference probes), sharded cyclically."""
    rank, world, dist = dist_info(group)
    thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    total, P = thetas.shape
    mine = list(range(rank, total, world))
    if want_grad and hasattr(model, "logPosterior_and_grad_batch") and len(mine) > 1:
        vals, grads = model.logPosterior_and_grad_batch([thetas[i] for i in mine], model.training_nodes, model.training_values)
        rows = [np.concatenate([[v], g]) for v, g in zip(vals, grads)]
    else:
        rows = []
        for i in mine:
            if want_grad and hasattr(model, "logPosterior_and_grad"):
                v, g = model.logPosterior_and_grad(thetas[i], model.training_nodes, model.training_values)
                rows.append(np.concatenate([[v], g]))
            else:
                v = model.logPosterior(thetas[i], model.training_nodes, model.training_values)
                rows.append(np.concatenate([[float(np.squeeze(v))], np.zeros(P)]))
    local = np.asarray(rows, dtype=np.float64).reshape(-1, 1 + P)
    return allgather_rows(local, total, rank, world, dist, group)
