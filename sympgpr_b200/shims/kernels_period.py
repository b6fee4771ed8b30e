"""Drop-in for the f2py extension `kernels_period` (python -m numpy.f2py -m kernels_period -c kernels_period.f90,
python/04_standard_map/Makefile:6-7): 19 scalar functions f(x_a, y_a, x_b, y_b, lx, ly, p) -> float."""
from sympgpr_b200.api import SCALAR_NAMES as _NAMES, scalar_function as _f

for _n in _NAMES:
    globals()[_n] = _f("product", _n, True)
__all__ = list(_NAMES)
