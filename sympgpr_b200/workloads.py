"""Deterministic synthetic workloads for benchmarks and examples (host side, NumPy).

The standard-map training set follows python/04_standard_map/main.py:27-59,89-92 with the
stochasticity parameter k -> 0.9 and unscrambled Halton points (bases 2,3; indices 1..N) on
[0, 2pi)^2, as fixed in SURVEY.md section 8(d) / BASELINE.md."""
import numpy as np

TWO_PI = 2.0 * np.pi


def halton(n, base, start=1):
    """Van der Corput sequence in `base`, indices start..start+n-1 (vectorised)."""
    idx = np.arange(start, start + n, dtype=np.int64)
    out = np.zeros(n)
    f = 1.0
    while np.any(idx > 0):
        f /= base
        out += f * (idx % base)
        idx //= base
    return out


def standard_map_training(N, kchaos=0.9):
    q = halton(N, 2) * TWO_PI
    p = halton(N, 3) * TWO_PI
    P = p + kchaos * np.sin(q)
    Q = q + P
    xtrain = np.hstack((q, P))
    ztrain = np.concatenate((p - P, Q - q))
    xtrainp = np.hstack((q, p))
    ztrainp = P - p
    return dict(q=q, p=p, Q=Q, P=P, xtrain=xtrain, ztrain=ztrain, xtrainp=xtrainp, ztrainp=ztrainp,
                sig=2 * np.amax(np.abs(ztrain))**2, sigp=2 * np.amax(np.abs(ztrainp))**2)


def timing_hyp(N, sig, sig2n=1e-8, factor=0.5):
    """lx = ly = factor * 2pi / sqrt(N); factor 0.5 keeps cond(Ky) ~ 1e4."""
    l = factor * TWO_PI / np.sqrt(N)
    return np.array([l, l, sig, sig2n])


def ensemble(E, lo=1.0, hi=5.0):
    """Initial conditions: Halton bases 5, 7; q in [0, 2pi), p in [lo, hi)."""
    return halton(E, 5) * TWO_PI, lo + halton(E, 7) * (hi - lo)


def henon_like_training(N, seed=3):
    """BASELINE config 3 (2-DOF, 4 x 4-block kernel; not in the reference): one kick-drift step of a Henon-Heiles-like
    potential V = (q1^2 + q2^2)/2 + q1^2 q2 - q2^3/3 with dt = 0.3, P = p - dt dV/dq(q), Q = q + dt P, on Halton points in
    [-0.4, 0.4]^4.  Returns x = [q1; q2; P1; P2] (4N) and z = [p1 - P1; p2 - P2; Q1 - q1; Q2 - q2] (4N)."""
    dt = 0.3
    q1 = -0.4 + 0.8 * halton(N, 2, seed)
    q2 = -0.4 + 0.8 * halton(N, 3, seed)
    p1 = -0.4 + 0.8 * halton(N, 5, seed)
    p2 = -0.4 + 0.8 * halton(N, 7, seed)
    P1 = p1 - dt * (q1 + 2 * q1 * q2)
    P2 = p2 - dt * (q2 + q1**2 - q2**2)
    Q1, Q2 = q1 + dt * P1, q2 + dt * P2
    return np.concatenate((q1, q2, P1, P2)), np.concatenate((p1 - P1, p2 - P2, Q1 - q1, Q2 - q2))


def dof2_hyp(N, z, shrink=1.0):
    """Length scales of the 2-DOF benchmark model (tools/bench_dof2.py of round 1): ~ N^(-1/4) in four dimensions."""
    s = shrink * (200.0 / N) ** 0.25
    return np.array([0.35 * s, 0.4 * s, 2 * np.max(np.abs(z))**2, 1e-6])


def pendulum_training(N, U0=1.0, dt=0.5, pmax=2.5):
    """BASELINE config 1 (01_pendulum, ~200 training pairs): one kick-drift step of the pendulum H = p^2/2 + U0 (1 - cos q),
    P = p - dt U0 sin q, Q = q + dt P, on Halton points q in [0, 2pi), p in [-pmax, pmax).  Layout as
    python/01_pendulum/implicit/main.py:110-125: xtrain = [q; P], ztrain = [p - P; Q - q], xtrainp = [q; p], ztrainp = P
    (the pendulum scripts train the guess GP on P itself, so the reference's own hybrd1 start is the consistent one)."""
    q = halton(N, 2) * TWO_PI
    p = -pmax + 2 * pmax * halton(N, 3)
    P = p - dt * U0 * np.sin(q)
    Q = q + dt * P
    xtrain, ztrain = np.hstack((q, P)), np.concatenate((p - P, Q - q))
    xtrainp, ztrainp = np.hstack((q, p)), P.copy()
    return dict(q=q, p=p, Q=Q, P=P, xtrain=xtrain, ztrain=ztrain, xtrainp=xtrainp, ztrainp=ztrainp,
                sig=2 * np.amax(np.abs(ztrain))**2, sigp=2 * np.amax(np.abs(ztrainp))**2)


def tokamak_training(N, eps=0.08, mpol=2, iota0=0.3, shear=0.06, plo=0.3, phi=9.7):
    """BASELINE config 5 (05_tokamak): a field-line-like twist map in the reference's map coordinates (theta, p = 1e2 p_theta;
    python/05_tokamak/SympGPR/calc_fieldlines.py scales p_theta by 1e2 before training):
        P = p - eps mpol sin(mpol theta),   Theta = theta + 2 pi (iota0 + shear P) / 8
    -- one eighth of a toroidal turn per map step with a rotational transform that grows with the flux label and an
    (mpol, n) island chain; a closed-form stand-in for fieldlines.timestep (the integrator is a data
    generator outside the hot path).  Orbits are lost where compute_r([1e-2 P, theta, 0], 0.3) > 0.5 or P < 0
    (python/05_tokamak/SympGPR/func.py:194-203), i.e. beyond P ~ 8.3 ... 16.7 depending on theta.
    Guess GP trained on P - p as python/05_tokamak/SympGPR/main.py:34-35 does."""
    th = halton(N, 2) * TWO_PI
    p = plo + (phi - plo) * halton(N, 3)
    P = p - eps * mpol * np.sin(mpol * th)
    Th = th + TWO_PI * (iota0 + shear * P) / 8.0
    xtrain, ztrain = np.hstack((th, P)), np.concatenate((p - P, Th - th))
    xtrainp, ztrainp = np.hstack((th, p)), P - p
    return dict(q=th, p=p, Q=Th, P=P, xtrain=xtrain, ztrain=ztrain, xtrainp=xtrainp, ztrainp=ztrainp,
                sig=2 * np.amax(np.abs(ztrain))**2, sigp=2 * np.amax(np.abs(ztrainp))**2,
                par=dict(eps=eps, mpol=mpol, iota0=iota0, shear=shear))


def tokamak_exact_step(th, p, par):
    """The map tokamak_training() samples, for sanity checks of the learned map."""
    P = p - par["eps"] * par["mpol"] * np.sin(par["mpol"] * th)
    return np.mod(th + TWO_PI * (par["iota0"] + par["shear"] * P) / 8.0, TWO_PI), P


def aniso_hyp(N, sig, lq_span=TWO_PI, lp_span=TWO_PI, factor=1.0, sig2n=1e-8):
    """lx, ly = factor * span / sqrt(N) per axis (timing_hyp for a domain that is not square)."""
    return np.array([factor * lq_span / np.sqrt(N), factor * lp_span / np.sqrt(N), sig, sig2n])
