"""Drop-in for the f2py extension `sympgpr` (python/05_tokamak/SympGPR/make_sympgpr.mk):
`from sympgpr import sympgpr` yields an object with the Fortran module's callables
(signatures: SURVEY.md Appendix D), backed by libsympgpr_b200.so on the GPU."""
import numpy as _np

from sympgpr_b200 import api as _api


class _SympgprModule:
    """Stands in for the f2py Fortran-module object `sympgpr.sympgpr`."""

    pi = 4.0 * _np.arctan(1.0)                    # sympgpr.f90:8

    # kernel family / period / root solver used by the calls below; the reference selects the
    # family by which kernels*.f90 was linked into the extension
    family = "product"
    per = 0.5
    solver = "hybrd"

    def build_k(self, x, y, x0, y0, hyp, k):
        _api.build_k(x, y, x0, y0, hyp, k, self.family, self.per)

    def buildkreg(self, x, y, x0, y0, hyp, k):
        _api.buildkreg(x, y, x0, y0, hyp, k, self.family, self.per)

    def guessp(self, x, y, hypp, xtrainp, ytrainp, ztrainp, kyinvp):
        return _api.guessp(x, y, hypp, xtrainp, ytrainp, ztrainp, kyinvp, self.family, self.per)

    def calcq(self, x, y, xtrain, ytrain, hyp, kyinv, ztrain):
        return _api.calcq(x, y, xtrain, ytrain, hyp, kyinv, ztrain, self.family, self.per)

    def calcp(self, x, y, hyp, hypp, xtrainp, ytrainp, ztrainp, kyinvp, xtrain, ytrain, ztrain, kyinv):
        return _api.calcp(x, y, hyp, hypp, xtrainp, ytrainp, ztrainp, kyinvp, xtrain, ytrain, ztrain, kyinv,
                          self.family, self.per, self.solver)

    def applymap_tok(self, hyp, hypp, q0map, p0map, xtrainp, ytrainp, ztrainp, kyinvp, xtrain, ytrain, ztrain, kyinv,
                     qmap, pmap, nm=None, ntest=None):
        _api.applymap_tok_f2py(hyp, hypp, q0map, p0map, xtrainp, ytrainp, ztrainp, kyinvp, xtrain, ytrain, ztrain, kyinv,
                               qmap, pmap, nm, ntest, self.family, self.per, self.solver)


sympgpr = _SympgprModule()
