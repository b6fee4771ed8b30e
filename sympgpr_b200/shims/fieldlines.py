"""Drop-in for the hot-path part of the f2py extension `fieldlines`
(python/05_tokamak/SympGPR/make_fieldlines.mk): `from fieldlines import fieldlines`.
Only what the map's loss test touches is provided (compute_r, ath; python/05_tokamak/SympGPR/
func.py:200-203); timestep/init are training-data generation and out of scope."""
import numpy as _np

from sympgpr_b200 import api as _api


class _Fieldlines:
    pi = 4.0 * _np.arctan(1.0)
    dph = 0.0
    eps = 0.0
    phase = 0.0
    m = 0
    n = 0
    rlast = 0.0

    def init(self, nph, am, an, aeps, aphase, arlast):          # fieldlines.f90:21-31
        self.dph = 2 * self.pi / nph
        self.m, self.n, self.eps, self.phase, self.rlast = am, an, aeps, aphase, arlast

    def compute_r(self, z, rstart):
        return _api.compute_r(z, rstart)

    def ath(self, r, th, ph):
        return _api.ath(r, th, ph)

    def timestep(self, z):
        raise NotImplementedError("fieldlines.timestep is training-data generation (out of scope, DESIGN.md)")


fieldlines = _Fieldlines()
