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


def _agree_or_raise(err, what, group=None, device=None):
    """Every rank calls this after its local work with err = None or the exception it caught.  One all_reduce(MAX) of a
    flag; if any rank failed, EVERY rank raises -- so no rank is left waiting in the collective that follows while a
    peer has already left the function with an exception."""
    import torch
    d = _dist()
    rank, world = rank_world(group)
    flag = 0 if err is None else 1
    if d is not None and world > 1:
        t = torch.tensor([flag], dtype=torch.int32)
        if device is not None:
            t = t.to(device)
        d.all_reduce(t, op=d.ReduceOp.MAX, group=group)
        flag = int(t.item())
    if err is not None:
        raise err
    if flag:
        raise RuntimeError(f"{what}: another rank failed (see its traceback); this rank stopped before the collective")


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
    err, q, p = None, None, None
    try:
        q, p = step_fn(Q0map[idx], P0map[idx])
    except Exception as e:                      # noqa: BLE001 -- re-raised on every rank below
        err = e
    _agree_or_raise(err, "applymap_sharded", group, device)
    return gather_interleaved(q, E, group, device), gather_interleaved(p, E, group, device)


def restarts_sharded(thetas, evaluate, group=None, device=None):
    """Evaluate `evaluate(theta) -> (value, grad)` for a list of hyper-parameter vectors, one per
    rank at a time; returns (values (T,), grads (T, G)) on every rank.

    A start whose kernel matrix is not positive definite (numpy.linalg.LinAlgError from api.nll_*, the expected outcome
    of a bad random start -- the reference scripts wrap their NLL in try/except for it, python/02_pert_pendulum/func.py:
    194-204) or whose arguments are rejected (ValueError) is recorded as value = +inf, grad = NaN and the sweep goes on.
    Any other exception is re-raised on EVERY rank after a flag exchange, so no rank hangs in the gather."""
    import torch
    d = _dist()
    rank, world = rank_world(group)
    thetas = [np.asarray(t, float) for t in thetas]
    T = len(thetas)
    mine = list(range(rank, T, world))
    res, err = [], None
    for i in mine:
        try:
            res.append(evaluate(thetas[i]))
        except (np.linalg.LinAlgError, ValueError):
            res.append((np.inf, None))
        except Exception as e:                  # noqa: BLE001 -- re-raised on every rank below
            err = e
            break
    _agree_or_raise(err, "restarts_sharded", group, device)
    G = max([len(np.atleast_1d(r[1])) for r in res if r[1] is not None] + [0])
    res = [(r[0], np.full(G, np.nan) if r[1] is None else r[1]) for r in res]
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
        g = np.atleast_1d(r[1])
        buf[j, 1:1 + len(g)] = g
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
    err, out = None, None
    try:
        out = quality_fn(Q0map[idx], P0map[idx])
    except Exception as e:                      # noqa: BLE001 -- re-raised on every rank below
        err = e
    _agree_or_raise(err, "quality_sharded", group, device)
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


def _radical_inverse_t(idx, base):
    """torch twin of _radical_inverse: idx int64 tensor on any device."""
    import torch
    out = torch.zeros(idx.shape, dtype=torch.float64, device=idx.device)
    f = 1.0
    idx = idx.clone()
    while bool((idx > 0).any()):
        f /= base
        out += f * (idx % base).to(torch.float64)
        idx = idx // base
    return out


def saltelli_block_t(start, count, bounds, device):
    """saltelli_block with the rows generated on `device` (torch): no host arrays, no H2D copy of the sample set."""
    import torch
    d = len(bounds)
    if 2 * d > len(_PRIMES):
        raise ValueError("saltelli_block: at most %d inputs" % (len(_PRIMES) // 2))
    idx = torch.arange(start + 1, start + 1 + count, dtype=torch.int64, device=device)
    lo = torch.tensor([b[0] for b in bounds], dtype=torch.float64, device=device)
    hi = torch.tensor([b[1] for b in bounds], dtype=torch.float64, device=device)
    A = torch.stack([_radical_inverse_t(idx, _PRIMES[k]) for k in range(d)], dim=1) * (hi - lo) + lo
    B = torch.stack([_radical_inverse_t(idx, _PRIMES[d + k]) for k in range(d)], dim=1) * (hi - lo) + lo
    return A, B


def sobol_indices_sharded(model, bounds, n_samples, group=None, device=None, block=1 << 18, on_device=False):
    """First-order and total Sobol indices of a scalar model output over a sample set split across the ranks.

    model(X) -> y maps an (m, d) block of inputs to m outputs (e.g. a closure that applies the learned map to
    initial conditions X[:, 0], X[:, 1] for S steps and returns the final action); NaN outputs (lost orbits) are
    dropped from every estimator together with their row.
    Rank r evaluates the contiguous row range [r n/world, (r+1) n/world) of the Saltelli matrices: d + 2 model
    runs per row (A, B and the d matrices A with column i taken from B).  Only the estimator sums are
    communicated: ONE all_reduce(sum) of 4 + 3 d doubles at the end (NCCL on GPUs, gloo on CPU).

    on_device=True: the sample rows are generated on `device` (torch), `model` takes and returns torch tensors on
    that device (map_model_on_device() below builds such a closure over the C ABI's device-pointer entry points), the
    estimator sums are accumulated there and the all_reduce runs on the device tensor -- nothing but the final
    4 + 3 d numbers ever reaches the host.  on_device=False: NumPy arrays on the host (CPU tests, small sets).

    Estimators (Saltelli 2010 / Jansen 1999):  V = var(f(A) u f(B)),
        S_i  = mean( f(B) (f(AB_i) - f(A)) ) / V,      ST_i = mean( (f(A) - f(AB_i))^2 ) / (2 V).
    Returns dict(S1, ST, mean, var, n_used) -- identical on every rank."""
    import torch
    dd = _dist()
    rank, world = rank_world(group)
    d = len(bounds)
    n = int(n_samples)
    r0, r1 = rank * n // world, (rank + 1) * n // world
    # [count, sum y, sum y^2 (over A and B), spare, then per i: sum fB(fABi-fA), sum (fA-fABi)^2, count_i]
    err = None
    if on_device:
        acc = torch.zeros(4 + 3 * d, dtype=torch.float64, device=device)
    else:
        acc = np.zeros(4 + 3 * d)
    try:
        for s in range(r0, r1, block):
            m = min(block, r1 - s)
            if on_device:
                A, B = saltelli_block_t(s, m, bounds, device)
                fA, fB = model(A), model(B)
                ok = torch.isfinite(fA) & torch.isfinite(fB)
                zA, zB = torch.where(ok, fA, torch.zeros_like(fA)), torch.where(ok, fB, torch.zeros_like(fB))
                acc[0] += 2 * ok.sum()
                acc[1] += zA.sum() + zB.sum()
                acc[2] += (zA * zA).sum() + (zB * zB).sum()
                for i in range(d):
                    AB = A.clone()
                    AB[:, i] = B[:, i]
                    fAB = model(AB)
                    oki = ok & torch.isfinite(fAB)
                    z = torch.zeros_like(fAB)
                    a_, b_, ab_ = torch.where(oki, fA, z), torch.where(oki, fB, z), torch.where(oki, fAB, z)
                    acc[4 + 3 * i] += (b_ * (ab_ - a_)).sum()
                    acc[5 + 3 * i] += ((a_ - ab_)**2).sum()
                    acc[6 + 3 * i] += oki.sum()
            else:
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
    except Exception as e:                      # noqa: BLE001 -- re-raised on every rank below
        err = e
    _agree_or_raise(err, "sobol_indices_sharded", group, device)
    if dd is not None and world > 1:
        t = acc if on_device else torch.from_numpy(acc)
        if device is not None and not on_device:
            t = t.to(device)
        dd.all_reduce(t, op=dd.ReduceOp.SUM, group=group)          # the only collective of the path
        acc = t
    acc = acc.cpu().numpy() if hasattr(acc, "cpu") else acc
    cnt = max(acc[0], 1.0)
    mean = acc[1] / cnt
    var = acc[2] / cnt - mean**2
    S1 = np.array([acc[4 + 3 * i] / max(acc[6 + 3 * i], 1.0) for i in range(d)]) / var
    ST = np.array([acc[5 + 3 * i] / max(acc[6 + 3 * i], 1.0) for i in range(d)]) / (2.0 * var)
    return dict(S1=S1, ST=ST, mean=float(mean), var=float(var), n_used=int(acc[0] // 2),
                n_total=int(n), model_runs=int(n) * (d + 2))


class DeviceMapModel:
    """A learned map resident on one GPU (sgp_model_create) with torch-tensor entry points: the closure that
    sobol_indices_sharded(on_device=True), applymap_sharded and bench.py apply to ensembles that never leave the device.

        m = DeviceMapModel(hyp3, hypp3, xtrainp, alphap, xtrain, alpha, family="product")
        qf, pf = m.applymap(q0, p0, nsteps, kind="tokamak", solver="newton_delta")     # torch tensors on m.device

    q0, p0: float64 CUDA tensors (E,).  Returns the final states; lost orbits are NaN.  Calls enqueue on the context's
    stream, which is made to wait for / be waited on by torch's current stream."""

    def __init__(self, hyp3, hypp3, xtrainp, alphap, xtrain, alpha, family="product", per=0.5, device=None):
        import ctypes
        import torch
        from . import _lib
        self._lib, self._torch = _lib, torch
        self.device = torch.device("cuda", _lib.default_device()) if device is None else torch.device(device)
        self.ctx = _lib.context(self.device.index)
        xtrainp, xtrain = _lib.as_f64(xtrainp).ravel(), _lib.as_f64(xtrain).ravel()
        alphap, alpha = _lib.as_f64(alphap).ravel(), _lib.as_f64(alpha).ravel()
        np_, nt = xtrainp.size // 2, xtrain.size // 2
        if alphap.size != np_ or alpha.size != 2 * nt:
            raise ValueError("DeviceMapModel: alpha vectors do not match the training sets")
        self.np_, self.nt = np_, nt
        self.family = _lib.FAMILIES[family] if isinstance(family, str) else int(family)
        h = ctypes.c_void_p()
        dp = _lib.dptr
        _lib.check(_lib.lib().sgp_model_create(
            self.ctx.handle, self.family, float(per), dp(_lib.as_f64(hyp3).ravel()[:3].copy()), dp(_lib.as_f64(hypp3).ravel()[:3].copy()),
            dp(np.ascontiguousarray(xtrainp[:np_])), dp(np.ascontiguousarray(xtrainp[np_:])), dp(alphap), np_,
            dp(np.ascontiguousarray(xtrain[:nt])), dp(np.ascontiguousarray(xtrain[nt:])), dp(alpha), nt, ctypes.byref(h)),
            "sgp_model_create")
        self.handle = h
        self.stats = torch.zeros(2, dtype=torch.int64, device=self.device)

    def applymap(self, q0, p0, nsteps, kind="pendulum", solver="hybrd"):
        torch, _lib = self._torch, self._lib
        if q0.dtype != torch.float64 or p0.dtype != torch.float64 or q0.device != self.device or p0.device != self.device:
            raise ValueError("DeviceMapModel.applymap: q0, p0 must be float64 tensors on the model's device")
        q0, p0 = q0.contiguous(), p0.contiguous()
        E = int(q0.numel())
        qf, pf = torch.empty_like(q0), torch.empty_like(p0)
        if E == 0:
            return qf, pf
        k = _lib.MAP_KINDS[kind] if isinstance(kind, str) else int(kind)
        sv = _lib.SOLVERS[solver] if isinstance(solver, str) else int(solver)
        cur = torch.cuda.current_stream(self.device)
        self.ctx.set_stream(cur.cuda_stream if cur.cuda_stream != 0 else None)     # same stream as the producer of q0/p0
        if cur.cuda_stream == 0:
            torch.cuda.synchronize(self.device)                                    # legacy default stream: order by a sync
        _lib.check(_lib.lib().sgp_model_applymap_dev(self.ctx.handle, self.handle, k, sv, int(nsteps), E, q0.data_ptr(),
                                                     p0.data_ptr(), qf.data_ptr(), pf.data_ptr(), None, None, 0,
                                                     self.stats.data_ptr()), "sgp_model_applymap_dev")
        if cur.cuda_stream == 0:
            self.ctx.synchronize()
        return qf, pf

    def close(self):
        if getattr(self, "handle", None):
            self._lib.lib().sgp_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
