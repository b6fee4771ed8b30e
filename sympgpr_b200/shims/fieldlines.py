"""Drop-in for the f2py extension `fieldlines` (python/05_tokamak/SympGPR/make_fieldlines.mk):
`from fieldlines import fieldlines`.

Hot-path part (the map's loss test, python/05_tokamak/SympGPR/func.py:200-203): compute_r, ath -- through the library.
The rest of the module (fieldlines.f90:21-172: init, the vector-potential closed forms, f_r, f_tstep, timestep) is the
generator of the training / reference field lines, which SURVEY 8(b) marks out of scope with a host implementation
acceptable "for running the script": plain Python below, statement by statement, with SciPy's MINPACK hybrd given
hybrd1's parameter block (minpack.f90:1570-1577) for the implicit midpoint step.  Module state (dph, eps, phase, m, n,
rlast) behaves like the Fortran module variables, including rlast being updated inside f_tstep (fieldlines.f90:121-122)."""
import math as _math

import numpy as _np

from sympgpr_b200 import api as _api


class _Fieldlines:
    pi = 4.0 * _np.arctan(1.0)
    B0 = 1.0
    iota0 = 1.0      # constant part of rotational transform
    a = 0.5          # (equivalent) minor radius
    R0 = 1.0         # (equivalent) major radius
    dph = 0.0
    eps = 0.0
    phase = 0.0
    m = 0
    n = 0
    rlast = 0.0

    def init(self, nph, am, an, aeps, aphase, arlast):          # fieldlines.f90:21-31
        self.dph = 2 * self.pi / nph
        self.m, self.n, self.eps, self.phase, self.rlast = am, an, aeps, aphase, arlast

    # ---- hot-path adjacent: through the library ------------------------------------------------
    def compute_r(self, z, rstart):                             # fieldlines.f90:94-107
        return _api.compute_r(z, rstart)

    def ath(self, r, th, ph):                                   # fieldlines.f90:34-39
        return _api.ath(r, th, ph)

    # ---- closed forms (fieldlines.f90:42-79) ---------------------------------------------------
    def dathdr(self, r, th, ph):
        return self.B0 * (r - r**2 / self.R0 * _math.cos(th))

    def dathdth(self, r, th, ph):
        return self.B0 * r**3 * _math.sin(th) / (3.0 * self.R0)

    def _pert(self, th, ph):
        return self.m * th + self.n * ph + self.phase

    def aph(self, r, th, ph):
        return -self.B0 * self.iota0 * (r**2 / 2.0 - r**4 / (4.0 * self.a**2)) * (1.0 + self.eps * _math.cos(self._pert(th, ph)))

    def daphdr(self, r, th, ph):
        return -self.B0 * self.iota0 * (r - r**3 / self.a**2) * (1.0 + self.eps * _math.cos(self._pert(th, ph)))

    def daphdth(self, r, th, ph):
        return self.B0 * self.iota0 * (r**2 / 2.0 - r**4 / (4.0 * self.a**2)) * self.m * self.eps * _math.sin(self._pert(th, ph))

    def f_r(self, x, args):                                     # fieldlines.f90:82-91 -> (y, dy)
        ct = _math.cos(args[1])
        return args[0] - self.B0 * (x**2 / 2.0 - x**3 / (3.0 * self.R0) * ct), -self.B0 * (x - x**2 / self.R0 * ct)

    def _compute_r(self, z, rstart):                            # host twin of compute_r for the generator (20 Newton steps)
        r = rstart
        for _ in range(20):
            y, dy = self.f_r(r, z)
            r = r - y / dy
        return r

    def f_tstep(self, znew, zold):                              # fieldlines.f90:110-140 -> (y(2), dy(2,2)); updates rlast
        z = [0.5 * (zold[0] + znew[0]), 0.5 * (zold[1] + znew[1]), zold[2] + 0.5 * self.dph]
        r = self._compute_r(z, self.rlast)
        self.rlast = r
        dApdr, dApdt = self.daphdr(r, z[1], z[2]), self.daphdth(r, z[1], z[2])
        dAtdr, dAtdt = self.dathdr(r, z[1], z[2]), self.dathdth(r, z[1], z[2])
        y = _np.array([zold[0] - znew[0] + self.dph * (dApdt - dApdr * dAtdt / dAtdr),
                       zold[1] - znew[1] - self.dph * dApdr / dAtdr])
        return y, _np.zeros((2, 2))                             # the Fortran leaves the Jacobian at zero ("TODO")

    def timestep(self, z):                                      # fieldlines.f90:143-170, in place on z(3)
        import scipy.optimize
        if not isinstance(z, _np.ndarray) or z.dtype != _np.float64 or z.shape != (3,):
            raise ValueError("failed to initialize intent(inout) array -- expected a float64 array of shape (3,)")
        zold = z.copy()
        sol = scipy.optimize.fsolve(lambda x: self.f_tstep(x, zold)[0], z[:2].copy(), xtol=1e-13, maxfev=600,
                                    diag=[1.0, 1.0], factor=100, epsfcn=0.0)
        z[0], z[1] = sol[0], sol[1]
        z[2] = zold[2] + self.dph


fieldlines = _Fieldlines()
