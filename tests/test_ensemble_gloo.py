"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: sharding an orbit ensemble / a list
of restarts over ranks and gathering the results.  The per-rank compute is a deterministic NumPy
stand-in here (the CUDA kernels need a GPU); the GPU-side use is bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_map(q0, p0, nm=4):
    """stand-in for api.applymap: a standard-map history (nm, E_local)"""
    q, p = np.zeros((nm, len(q0))), np.zeros((nm, len(q0)))
    q[0], p[0] = q0, p0
    for i in range(nm - 1):
        p[i + 1] = p[i] + 0.9 * np.sin(q[i])
        q[i + 1] = np.mod(q[i] + p[i + 1], 2 * np.pi)
    return q, p


def _worker(rank, world, port, E, T, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sympgpr_b200 import ensemble as ens
    rng = np.random.default_rng(0)
    q0, p0 = rng.uniform(0, 6.28, E), rng.uniform(0, 6.28, E)
    q, p = ens.applymap_sharded("standard", 4, q0, p0, _fake_map)
    qr, pr = _fake_map(q0, p0)
    ok1 = np.array_equal(q, qr) and np.array_equal(p, pr)
    idx = ens.shard_indices(E, rank, world)
    ok2 = len(idx) in (E // world, E // world + 1) and (len(idx) == 0 or idx[0] == rank)
    thetas = [np.array([0.1 * (i + 1), 0.2 * (i + 1)]) for i in range(T)]
    calls = []

    def evaluate(th):
        calls.append(th)
        return float(np.sum(th**2)), 2 * th
    v, g = ens.restarts_sharded(thetas, evaluate)
    ok3 = np.allclose(v, [np.sum(t**2) for t in thetas]) and np.allclose(g, [2 * t for t in thetas])
    ok4 = len(calls) == len(range(rank, T, world))
    out[rank] = int(ok1 and ok2 and ok3 and ok4)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("E,T", [(11, 5), (8, 2), (1, 1)])
def test_sharded_ensemble_and_restarts_world2(E, T):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, E, T, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1] * world


def _ishigami(X, a=7.0, b=0.1):
    return np.sin(X[:, 0]) + a * np.sin(X[:, 1])**2 + b * X[:, 2]**4 * np.sin(X[:, 0])


def _sobol_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sympgpr_b200 import ensemble as ens
    calls = []

    def model(X):
        calls.append(len(X))
        y = _ishigami(X)
        y[(X[:, 0] > 3.0) & (X[:, 1] > 3.0)] = np.nan          # "lost orbits" are dropped consistently
        return y
    res = ens.sobol_indices_sharded(model, [(-np.pi, np.pi)] * 3, 20001, block=4096)
    for k, v in (("S1", res["S1"]), ("ST", res["ST"])):
        out[rank * 8 + (0 if k == "S1" else 3):rank * 8 + (3 if k == "S1" else 6)] = list(v)
    out[rank * 8 + 6] = res["var"]
    out[rank * 8 + 7] = float(sum(calls))
    dist.barrier()
    dist.destroy_process_group()


def test_sobol_sample_sets_world2():
    """sample sets split over the ranks, one all-reduce of the estimator sums."""
    from sympgpr_b200 import ensemble as ens
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("d", [0.0] * (8 * world))
    procs = [ctx.Process(target=_sobol_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    o = np.array(out[:]).reshape(world, 8)
    assert np.array_equal(o[0, :7], o[1, :7])                   # every rank holds the same result
    assert o[0, 7] + o[1, 7] == 20001 * 5                       # rows split, d + 2 runs per row
    # the same computation in one process

    def model(X):
        y = _ishigami(X)
        y[(X[:, 0] > 3.0) & (X[:, 1] > 3.0)] = np.nan
        return y
    ref = ens.sobol_indices_sharded(model, [(-np.pi, np.pi)] * 3, 20001, block=100000)
    assert np.allclose(o[0, :3], ref["S1"], rtol=1e-10, atol=1e-12) and np.allclose(o[0, 3:6], ref["ST"], rtol=1e-10, atol=1e-12)
    # analytic Ishigami indices (a = 7, b = 0.1): S = (0.3139, 0.4424, 0), ST = (0.5576, 0.4424, 0.2437)
    clean = ens.sobol_indices_sharded(_ishigami, [(-np.pi, np.pi)] * 3, 60000)
    assert np.allclose(clean["S1"], [0.3139, 0.4424, 0.0], atol=0.02)
    assert np.allclose(clean["ST"], [0.5576, 0.4424, 0.2437], atol=0.02)


def test_single_process_paths():
    from sympgpr_b200 import ensemble as ens
    q0, p0 = np.linspace(0, 6, 7), np.linspace(1, 2, 7)
    q, p = ens.applymap_sharded("standard", 4, q0, p0, _fake_map)
    qr, pr = _fake_map(q0, p0)
    assert np.array_equal(q, qr) and np.array_equal(p, pr)
    v, g = ens.restarts_sharded([np.array([1.0, 2.0])], lambda t: (t.sum(), t))
    assert v.shape == (1,) and g.shape == (1, 2)


def _fake_quality(q0, p0):
    """stand-in for api.applymap_quality: first mapped state of the standard map and a made-up energy oscillation;
    orbits with p0 < 0.3 are 'lost' (NaN)"""
    P = p0 + 0.9 * np.sin(q0)
    q1, p1 = np.mod(q0 + P, 2 * np.pi), P.copy()
    eo = 1e-3 * (1 + np.cos(q0))
    lost = p0 < 0.3
    q1[lost] = np.nan; p1[lost] = np.nan; eo[lost] = np.nan
    return dict(q1=q1, p1=p1, Eosc=eo)


def _quality_reference(q0, p0, ys, Nm):
    o = _fake_quality(q0, p0)
    gd = np.mean((np.array([o["q1"], o["p1"]]) - ys[Nm, :2])**2, axis=0)
    ok = np.isfinite(gd) & np.isfinite(o["Eosc"])
    return dict(stdgd=np.std(gd[ok]), gd_mean=gd[ok].mean(), Eosc_mean=o["Eosc"][ok].mean(), Eosc_max=o["Eosc"][ok].max(),
                n=int(ok.sum()), n_lost=int((~ok).sum()))


def _quality_worker(rank, world, port, E, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sympgpr_b200 import ensemble as ens
    rng = np.random.default_rng(5)
    q0, p0 = rng.uniform(0, 6.28, E), rng.uniform(0, 6.28, E)
    ys = rng.uniform(0, 6.28, (3, 2, E))
    res = ens.quality_sharded(q0, p0, ys, 2, _fake_quality)
    ref = _quality_reference(q0, p0, ys, 2)
    ok = all(np.isclose(res[k], ref[k], rtol=1e-12, atol=0) for k in ("stdgd", "gd_mean", "Eosc_mean", "Eosc_max"))
    ok = ok and res["n"] == ref["n"] and res["n_lost"] == ref["n_lost"] and res["n_lost"] > 0
    out[rank] = 1 if ok else 0
    dist.destroy_process_group()


def test_quality_statistics_world2():
    """`quality` of the reference (functions/func.py:262-272) over an ensemble split on 2 ranks: gd statistics and energy
    oscillation agree with the single-process evaluation on every rank; lost (NaN) orbits are counted, not averaged."""
    world, E = 2, 101
    port = _free_port()
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_quality_worker, args=(world, port, E, out), nprocs=world, join=True)
    assert all(out[r] == 1 for r in range(world))
    # single-process path (no process group)
    from sympgpr_b200 import ensemble as ens
    rng = np.random.default_rng(5)
    q0, p0 = rng.uniform(0, 6.28, E), rng.uniform(0, 6.28, E)
    ys = rng.uniform(0, 6.28, (3, 2, E))
    res = ens.quality_sharded(q0, p0, ys, 2, _fake_quality)
    ref = _quality_reference(q0, p0, ys, 2)
    assert np.isclose(res["stdgd"], ref["stdgd"], rtol=1e-12) and res["n"] == ref["n"]


# ---------------------------------------------------------------------------------------------------
# failures on one rank must not leave the others waiting in a collective
# ---------------------------------------------------------------------------------------------------
def _failure_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sympgpr_b200 import ensemble as ens
    ok = True
    # (a) a start that is not positive definite on ONE rank: recorded as +inf / NaN, every rank gets all results
    thetas = [np.array([0.1 * (i + 1), 0.2 * (i + 1)]) for i in range(5)]

    def evaluate(th):
        if abs(th[0] - 0.2) < 1e-12:                   # theta 1 -> evaluated by rank 1
            raise np.linalg.LinAlgError("cholesky: 3-th leading minor of the array is not positive definite")
        return float(np.sum(th**2)), 2 * th
    v, g = ens.restarts_sharded(thetas, evaluate)
    ok &= bool(np.isinf(v[1]) and np.isnan(g[1]).all())
    ok &= bool(np.allclose(np.delete(v, 1), [np.sum(t**2) for i, t in enumerate(thetas) if i != 1]))
    ok &= bool(np.allclose(np.delete(g, 1, axis=0), [2 * t for i, t in enumerate(thetas) if i != 1]))
    # (b) an unexpected exception on ONE rank: EVERY rank raises, nobody hangs in the gather
    def bad_eval(th):
        if rank == 0:
            raise RuntimeError("device fell off the bus")
        return float(np.sum(th**2)), 2 * th
    try:
        ens.restarts_sharded(thetas, bad_eval)
        ok = False
    except RuntimeError as e:
        ok &= ("device fell off the bus" in str(e)) if rank == 0 else ("another rank failed" in str(e))
    # (c) same for an ensemble step and a Sobol model
    def bad_step(q0, p0):
        if rank == 1:
            raise ValueError("bad shard")
        return _fake_map(q0, p0)
    try:
        ens.applymap_sharded("standard", 4, np.linspace(0, 6, 9), np.linspace(1, 2, 9), bad_step)
        ok = False
    except (ValueError, RuntimeError) as e:
        ok &= ("bad shard" in str(e)) if rank == 1 else ("another rank failed" in str(e))

    def bad_model(X):
        if rank == 1:
            raise FloatingPointError("model blew up")
        return _ishigami(X)
    try:
        ens.sobol_indices_sharded(bad_model, [(-np.pi, np.pi)] * 3, 1024)
        ok = False
    except (FloatingPointError, RuntimeError) as e:
        ok &= ("model blew up" in str(e)) if rank == 1 else ("another rank failed" in str(e))
    # the group is still usable afterwards
    t = torch.ones(1)
    dist.all_reduce(t)
    ok &= bool(t.item() == world)
    out[rank] = int(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_one_rank_failing_does_not_hang_the_collectives_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_failure_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, "a rank hung or crashed"
    assert list(out) == [1] * world


def _sobol_torch_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sympgpr_b200 import ensemble as ens

    def ish_t(X, a=7.0, b=0.1):
        return torch.sin(X[:, 0]) + a * torch.sin(X[:, 1])**2 + b * X[:, 2]**4 * torch.sin(X[:, 0])
    bounds = [(-np.pi, np.pi)] * 3
    r = ens.sobol_indices_sharded(ish_t, bounds, 1 << 15, device="cpu", on_device=True, block=4096)
    rh = ens.sobol_indices_sharded(_ishigami, bounds, 1 << 15, block=4096)
    ok = np.allclose(r["S1"], rh["S1"], atol=1e-10) and np.allclose(r["ST"], rh["ST"], atol=1e-10) and r["n_used"] == 1 << 15
    out[rank] = int(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sobol_tensor_path_world2():
    """The on_device=True path of sobol_indices_sharded (rows generated with torch, sums accumulated in a tensor, the
    all_reduce on that tensor) with CPU tensors over gloo: same indices as the NumPy path, identical on both ranks."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_sobol_torch_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert list(out) == [1] * world
