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
