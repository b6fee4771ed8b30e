"""Multi-GPU host logic: only what shards naturally is sharded.

  * orbit ensembles       -- interleaved slices of (Q0map, P0map) per rank, model replicated
  * multi-start restarts  -- one hyper-parameter vector per rank at a time
  * quality statistics    -- `quality` of the reference per shard, one all_gather of six numbers
  * Sobol sample sets     -- Saltelli rows per rank, one all_reduce of the estimator sums
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


def quality_sharded(Q0map, P0map, ysint, Nm, quality_fn, order="qp", group=None, device=None):
    """The reference's `quality` (python/functions/func.py:262-272; tokamak variant Split_SympGPR/func.py:221-232) for an
    ensemble split over the ranks, with the statistics gathered by ONE all_reduce.

    quality_fn(q0_local, p0_local) -> dict with q1, p1, Eosc of the local orbits (a closure over
    sympgpr_b200.api.applymap_quality, whose map kernel accumulates Eosc without writing histories).  ysint (rows, >=2, E)
    holds the reference orbits of ALL E initial conditions; gd[k] is the mean squared distance of the first mapped state
    from ysint[Nm, :, k].  Returns dict(stdgd, gd_mean, Eosc_mean, Eosc_max, n) -- identical on every rank; orbits whose
    Eosc or gd is NaN (lost orbits) are left out of the sums and counted in n_lost.  numpy.std semantics (population)."""
    import torch
    d = _dist()
    rank, world = rank_world(group)
    Q0map, P0map = np.asarray(Q0map, float), np.asarray(P0map, float)
    ys = np.asarray(ysint, float)
    E = Q0map.shape[0]
    idx = shard_indices(E, rank, world)
    out = quality_fn(Q0map[idx], P0map[idx])
    q1, p1, eo = (np.asarray(out[k], float) for k in ("q1", "p1", "Eosc"))
    if order == "pq":
        ref = np.array([ys[Nm, 0, idx], np.mod(ys[Nm, 1, idx], 2 * np.pi)])
        mine = np.array([p1, q1])
    else:
        ref = np.array([ys[Nm, 0, idx], ys[Nm, 1, idx]])
        mine = np.array([q1, p1])
    gd = np.mean((mine - ref)**2, axis=0)
    ok = np.isfinite(gd) & np.isfinite(eo)
    acc = np.array([ok.sum(), gd[ok].sum(), (gd[ok]**2).sum(), eo[ok].sum(), (~ok).sum()], float)
    emax = eo[ok].max() if ok.any() else -np.inf
    if d is not None and world > 1:
        # the only collective: six numbers per rank (sums add, the maximum does not, hence a gather rather than a reduce)
        t = torch.from_numpy(np.append(acc, emax))
        if device is not None:
            t = t.to(device)
        parts = [torch.empty_like(t) for _ in range(world)]
        d.all_gather(parts, t, group=group)
        allv = np.stack([p_.cpu().numpy() for p_ in parts])
        acc, emax = allv[:, :5].sum(axis=0), allv[:, 5].max()
    n = max(acc[0], 1.0)
    gmean = acc[1] / n
    return dict(stdgd=float(np.sqrt(max(acc[2] / n - gmean**2, 0.0))), gd_mean=float(gmean), Eosc_mean=float(acc[3] / n),
                Eosc_max=float(emax), n=int(acc[0]), n_lost=int(acc[4]))


# --------------------------------------------------------------------------- Sobol sample sets
def _radical_inverse(idx, base):
    out = np.zeros(len(idx))
    f = 1.0
    idx = idx.copy()
    while np.any(idx > 0):
        f /= base
        out += f * (idx % base)
        idx //= base
    return out


_PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53)


def saltelli_block(start, count, bounds):
    """Rows start .. start+count-1 of the two Saltelli sample matrices A, B (count, d): unscrambled Halton
    points in 2 d dimensions (first d coordinates -> A, last d -> B), scaled to `bounds` = [(lo, hi)] * d.
    Any rank can generate any row range, so the sample set shards without communication."""
    d = len(bounds)
    if 2 * d > len(_PRIMES):
        raise ValueError("saltelli_block: at most %d inputs" % (len(_PRIMES) // 2))
    idx = np.arange(start + 1, start + 1 + count, dtype=np.int64)
    lo = np.array([b[0] for b in bounds], float)
    hi = np.array([b[1] for b in bounds], float)
    A = np.stack([_radical_inverse(idx, _PRIMES[k]) for k in range(d)], axis=1) * (hi - lo) + lo
    B = np.stack([_radical_inverse(idx, _PRIMES[d + k]) for k in range(d)], axis=1) * (hi - lo) + lo
    return A, B


def sobol_indices_sharded(model, bounds, n_samples, group=None, device=None, block=1 << 18):
    """First-order and total Sobol indices of a scalar model output over a sample set split across the ranks.

    model(X) -> y maps an (m, d) block of inputs to m outputs (e.g. a closure that applies the learned map to
    initial conditions X[:, 0], X[:, 1] for S steps with sympgpr_b200.api.applymap(..., out_every=0) and returns
    the final action); NaN outputs (lost orbits) are dropped from every estimator together with their row.
    Rank r evaluates the contiguous row range [r n/world, (r+1) n/world) of the Saltelli matrices: d + 2 model
    runs per row (A, B and the d matrices A with column i taken from B).  Only the estimator sums are
    communicated: ONE all_reduce(sum) of 4 + 3 d doubles at the end (NCCL on GPUs, gloo on CPU).

    Estimators (Saltelli 2010 / Jansen 1999):  V = var(f(A) u f(B)),
        S_i  = mean( f(B) (f(AB_i) - f(A)) ) / V,      ST_i = mean( (f(A) - f(AB_i))^2 ) / (2 V).
    Returns dict(S1, ST, mean, var, n_used) -- identical on every rank."""
    import torch
    dd = _dist()
    rank, world = rank_world(group)
    d = len(bounds)
    n = int(n_samples)
    r0, r1 = rank * n // world, (rank + 1) * n // world
    acc = np.zeros(4 + 3 * d)          # [count, sum y, sum y^2 (over A and B), spare, then per i: sum fB(fABi-fA), sum (fA-fABi)^2, count_i]
    for s in range(r0, r1, block):
        m = min(block, r1 - s)
        A, B = saltelli_block(s, m, bounds)
        fA, fB = np.asarray(model(A), float), np.asarray(model(B), float)
        ok = np.isfinite(fA) & np.isfinite(fB)
        acc[0] += 2 * ok.sum()
        acc[1] += fA[ok].sum() + fB[ok].sum()
        acc[2] += (fA[ok]**2).sum() + (fB[ok]**2).sum()
        for i in range(d):
            AB = A.copy()
            AB[:, i] = B[:, i]
            fAB = np.asarray(model(AB), float)
            oki = ok & np.isfinite(fAB)
            acc[4 + 3 * i] += (fB[oki] * (fAB[oki] - fA[oki])).sum()
            acc[5 + 3 * i] += ((fA[oki] - fAB[oki])**2).sum()
            acc[6 + 3 * i] += oki.sum()
    if dd is not None and world > 1:
        t = torch.from_numpy(acc)
        if device is not None:
            t = t.to(device)
        dd.all_reduce(t, op=dd.ReduceOp.SUM, group=group)
        acc = t.cpu().numpy()
    cnt = max(acc[0], 1.0)
    mean = acc[1] / cnt
    var = acc[2] / cnt - mean**2
    S1 = np.array([acc[4 + 3 * i] / max(acc[6 + 3 * i], 1.0) for i in range(d)]) / var
    ST = np.array([acc[5 + 3 * i] / max(acc[6 + 3 * i], 1.0) for i in range(d)]) / (2.0 * var)
    return dict(S1=S1, ST=ST, mean=mean, var=var, n_used=int(acc[0] // 2))
