"""Multi-GPU host logic: only what shards naturally is sharded.

  * orbit ensembles       -- interleaved slices of (Q0map, P0map) per rank, model replicated
  * multi-start restarts  -- one hyper-parameter vector per rank at a time
One process per GPU (torch.distributed, NCCL on GPUs / gloo in the CPU tests); there is no
collective on the data path, results are gathered once at the end.  The training Cholesky does
not shard (replicas only).
"""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def rank_world(group=None):
    d = _dist()
    if d is None:
        return 0, 1
    return d.get_rank(group), d.get_world_size(group)


def shard_indices(E, rank, world):
    """Interleaved assignment k = rank, rank+world, ...: lost/NaN orbits (which finish early) and
    slow-converging regions spread evenly over the ranks."""
    return np.arange(rank, E, world)


def gather_interleaved(local, E, group=None, device=None):
    """Inverse of shard_indices: every rank passes its slice (..., n_local) and receives the full
    (..., E) array.  Slices are padded to a common length for all_gather."""
    import torch
    d = _dist()
    rank, world = rank_world(group)
    local = np.asarray(local, dtype=np.float64)
    if d is None or world == 1:
        return local.copy()
    n_max = (E + world - 1) // world
    lead = local.shape[:-1]
    buf = np.full(lead + (n_max,), np.nan)
    buf[..., :local.shape[-1]] = local
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    d.all_gather(parts, t, group=group)
    out = np.empty(lead + (E,))
    for r in range(world):
        idx = shard_indices(E, r, world)
        out[..., idx] = parts[r].cpu().numpy()[..., :len(idx)]
    return out


def applymap_sharded(kind, nm, Q0map, P0map, step_fn, group=None, device=None):
    """Apply the map to an ensemble split over the ranks.

    step_fn(q0_local, p0_local) -> (q_local, p_local) with the ensemble on the last axis (e.g. a
    closure over sympgpr_b200.api.applymap with the model arguments bound); returns the gathered
    (q, p) for all E orbits on every rank."""
    Q0map, P0map = np.asarray(Q0map, float), np.asarray(P0map, float)
    E = Q0map.shape[0]
    rank, world = rank_world(group)
    idx = shard_indices(E, rank, world)
    q, p = step_fn(Q0map[idx], P0map[idx])
    return gather_interleaved(q, E, group, device), gather_interleaved(p, E, group, device)


def restarts_sharded(thetas, evaluate, group=None, device=None):
    """Evaluate `evaluate(theta) -> (value, grad)` for a list of hyper-parameter vectors, one per
    rank at a time; returns (values (T,), grads (T, G)) on every rank."""
    import torch
    d = _dist()
    rank, world = rank_world(group)
    thetas = [np.asarray(t, float) for t in thetas]
    T = len(thetas)
    mine = list(range(rank, T, world))
    res = [evaluate(thetas[i]) for i in mine]
    G = len(np.atleast_1d(res[0][1])) if res else 0
    if d is None or world == 1:
        return np.array([r[0] for r in res]), np.array([np.atleast_1d(r[1]) for r in res]).reshape(T, G)
    gt = torch.tensor([G], dtype=torch.int64)
    if device is not None:
        gt = gt.to(device)
    d.all_reduce(gt, op=d.ReduceOp.MAX, group=group)
    G = int(gt.item())
    n_max = (T + world - 1) // world
    buf = np.full((n_max, 1 + G), np.nan)
    for j, r in enumerate(res):
        buf[j, 0] = r[0]
        buf[j, 1:] = np.atleast_1d(r[1])
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    d.all_gather(parts, t, group=group)
    vals, grads = np.empty(T), np.empty((T, G))
    for r in range(world):
        a = parts[r].cpu().numpy()
        for j, i in enumerate(range(r, T, world)):
            vals[i], grads[i] = a[j, 0], a[j, 1:]
    return vals, grads
