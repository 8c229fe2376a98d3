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
    """Flat indices handled by ``rank``: rank, rank+world, ...  The grid's fast axis is the second one (noise),
    so consecutive flat indices share beta; striding by ``world`` gives every rank the same mix of Pade
    orders (cost depends on beta only)."""
    return rank, total, world


def allgather_rows(local: np.ndarray, total: int, rank: int, world: int, dist, group=None) -> np.ndarray:
    """Reassemble ``total`` rows from per-rank cyclic shards (rank r holds rows r, r+world, ...)."""
    width = local.shape[1]
    if world == 1 or dist is None:
        return local
    import torch
    per = (total + world - 1) // world
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    send = torch.zeros((per, width), dtype=torch.float64, device=device)
    send[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    recv = torch.empty((world * per, width), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, per, width)
    out = np.empty((total, width), dtype=np.float64)
    for r in range(world):
        cnt = len(range(r, total, world))
        out[r::world] = recv[r, :cnt]
    return out


def evaluate_slice(model, theta, axidx, ax1, ax2, begin, end, stride, want_grad=False):
    """(count, 1+P) array of (lml, dlml/dtheta) for the flat grid indices begin, begin+stride, ... < end."""
    from ._lib import KERNEL_IDS
    from .engine import raise_domain
    th = model._theta_vec(theta).copy()
    classifier = not hasattr(model, "logPosterior_and_grad")
    t = model.training_values
    t = np.asarray(getattr(t, "values", t), dtype=np.float64).reshape(-1)
    ctx = model._engine.bind(model._positions(model.training_nodes), model._positions(model.test_nodes))
    ctx._kernel_key = None
    ctx._predict_token = None
    st, out = ctx.landscape(KERNEL_IDS[model.kerneltype], classifier, th, axidx, ax1, ax2, t, begin, end, stride, want_grad)
    raise_domain(st)
    return out


def sharded_landscape(model, theta, axidx, ax1, ax2, want_grad=False, group=None) -> np.ndarray:
    """Full ``(len(ax1), len(ax2), 1+P)`` array on every rank: [..., 0] is the lml of GPnet.py:549,
    [..., 1:] its gradient when ``want_grad``."""
    rank, world, dist = dist_info(group)
    n1, n2 = len(ax1), len(ax2)
    total = n1 * n2
    begin, end, stride = cyclic_slice(total, rank, world)
    local = evaluate_slice(model, theta, axidx, ax1, ax2, begin, end, stride, want_grad)
    full = allgather_rows(local, total, rank, world, dist, group)
    return full.reshape(n1, n2, -1)


def sharded_multistart(model, thetas, want_grad=True, group=None) -> np.ndarray:
    """(-logp, -grad) for every row of ``thetas`` (multi-start / finite-difference probes), sharded cyclically."""
    rank, world, dist = dist_info(group)
    thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    total, P = thetas.shape
    rows = []
    for i in range(rank, total, world):
        if want_grad and hasattr(model, "logPosterior_and_grad"):
            v, g = model.logPosterior_and_grad(thetas[i], model.training_nodes, model.training_values)
            rows.append(np.concatenate([[v], g]))
        else:
            v = model.logPosterior(thetas[i], model.training_nodes, model.training_values)
            rows.append(np.concatenate([[float(np.squeeze(v))], np.zeros(P)]))
    local = np.asarray(rows, dtype=np.float64).reshape(-1, 1 + P)
    return allgather_rows(local, total, rank, world, dist, group)
