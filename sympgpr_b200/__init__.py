"""sympgpr_b200 -- B200-native hot path of SympGPR behind the reference's f2py module surface.

    from sympgpr_b200 import api            # build_K, nll_chol, nll_grad, applymap, ... (GPU)
    import sympgpr_b200; sympgpr_b200.install_shims()   # then `from sympgpr import sympgpr`,
                                                        # `from kernels import *`, ... resolve here
See DESIGN.md / INTEGRATION.md.  There is no CPU fallback: compute entry points raise when the
CUDA library or a CUDA device is missing.
"""
import os
import sys

__version__ = "0.1.0"

SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def install_shims(prepend_path=True):
    """Make `sympgpr`, `fortran.sympgpr`, `kernels`, `kernels_sq`, `kernels_sum`, `fieldlines`
    importable under the names the reference scripts use.  Modules are also pre-registered in
    sys.modules because the reference example directories contain stale `kernels.py` stubs that
    would otherwise win the path search ("import-path fidelity")."""
    import importlib

    if prepend_path and SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)
    names = ["sympgpr", "kernels", "kernels_sq", "kernels_sum", "kernels_period", "fieldlines", "fortran",
             "fortran.sympgpr"]
    mods = {}
    for name in names:
        spec_path = os.path.join(SHIM_DIR, *name.split(".")) + ".py"
        if not os.path.exists(spec_path):
            spec_path = os.path.join(SHIM_DIR, *name.split("."), "__init__.py")
        spec = importlib.util.spec_from_file_location(name, spec_path,
                                                      submodule_search_locations=[os.path.dirname(spec_path)]
                                                      if spec_path.endswith("__init__.py") else None)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    mods["fortran"].sympgpr = mods["fortran.sympgpr"]
    return mods
