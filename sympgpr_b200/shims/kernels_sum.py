"""Drop-in for the f2py extension `kernels_sum` (python -m numpy.f2py -m kernels_sum -c kernels_sum.f90,
python/04_standard_map/Makefile:6-7): 19 scalar functions f(x_a, y_a, x_b, y_b, lx, ly) -> float."""
from sympgpr_b200.api import SCALAR_NAMES as _NAMES, scalar_function as _f

for _n in _NAMES:
    globals()[_n] = _f("sum", _n, False)
__all__ = list(_NAMES)
