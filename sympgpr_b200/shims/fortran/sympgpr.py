"""`from fortran.sympgpr import sympgpr` (python/functions/func.py:13, python/02_pert_pendulum/func.py:13)."""
import importlib.util as _u
import os as _os
import sys as _sys

if "sympgpr" in _sys.modules and hasattr(_sys.modules["sympgpr"], "sympgpr"):
    sympgpr = _sys.modules["sympgpr"].sympgpr
else:
    _p = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "sympgpr.py")
    _spec = _u.spec_from_file_location("sympgpr", _p)
    _m = _u.module_from_spec(_spec)
    _spec.loader.exec_module(_m)
    sympgpr = _m.sympgpr
