"""ctypes binding of libsympgpr_b200.so (include/sympgpr_b200.h).

No CPU fallback: if the shared library is missing it is built with nvcc (sympgpr_b200/build.py);
if that is impossible, or no CUDA device is visible when a compute entry point is called, a
RuntimeError is raised.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SYMPGPR_B200_LIB: load another build of the same library (kernel experiments, tools/build_variant.py)
LIB_PATH = os.environ.get("SYMPGPR_B200_LIB") or os.path.join(_HERE, "libsympgpr_b200.so")

c_dp = ctypes.POINTER(ctypes.c_double)
c_ullp = ctypes.POINTER(ctypes.c_ulonglong)
c_vp = ctypes.c_void_p
c_d = ctypes.c_double
c_i = ctypes.c_int
c_l = ctypes.c_long

RES_LEN = 16
RES_NLL, RES_DLX, RES_DLY, RES_DSIG, RES_INFO, RES_QUAD, RES_LOGD, RES_A, RES_B = 0, 1, 2, 3, 4, 5, 6, 8, 11
E_BADARG, E_CUDA, E_NOMEM, E_NODEV = -1, -2, -3, -4

FAMILIES = {"product": 0, "sq": 1, "sum": 2, "period": 0}
MAP_KINDS = {"pendulum": 0, "henon": 1, "standard": 2, "tokamak": 3, "standard_expl": 4}
SOLVERS = {"hybrd": 0, "newton": 1, "explicit": 2, "newton_delta": 3}
ENERGIES = {"pendulum": 1, "tokamak": 2}

_lib = None
_lock = threading.Lock()
_ctx = {}

_PROTOS = {
    "sgp_version": (c_i, []),
    "sgp_launch_count": (ctypes.c_ulonglong, []),
    "sgp_last_error": (ctypes.c_char_p, []),
    "sgp_device_count": (c_i, []),
    "sgp_create": (c_i, [c_i, ctypes.POINTER(c_vp)]),
    "sgp_destroy": (c_i, [c_vp]),
    "sgp_set_stream": (c_i, [c_vp, c_vp]),
    "sgp_synchronize": (c_i, [c_vp]),
    "sgp_release_workspace": (c_i, [c_vp]),
    "sgp_set_profiling": (c_i, [c_vp, c_i]),
    "sgp_stage_times": (c_i, [c_vp, c_dp]),
    "sgp_kernel_scalar": (c_d, [c_i, c_i, c_d, c_d, c_d, c_d, c_d, c_d, c_d]),
    "sgp_build_k": (c_i, [c_vp, c_i, c_d, c_dp, c_dp, c_l, c_dp, c_dp, c_l, c_dp, c_dp, c_l]),
    "sgp_build_k4": (c_i, [c_vp, c_dp, c_l, c_dp, c_l, c_dp, c_dp, c_l]),
    "sgp_buildkreg": (c_i, [c_vp, c_i, c_d, c_dp, c_dp, c_l, c_dp, c_dp, c_l, c_dp, c_dp, c_l]),
    "sgp_guessp": (c_i, [c_vp, c_i, c_d, c_d, c_d, c_dp, c_dp, c_dp, c_dp, c_dp, c_l, c_dp]),
    "sgp_calcq": (c_i, [c_vp, c_i, c_d, c_d, c_d, c_dp, c_dp, c_dp, c_dp, c_dp, c_l, c_dp]),
    "sgp_calcp": (c_i, [c_vp, c_i, c_d, c_i, c_d, c_d, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_l, c_dp, c_dp, c_dp,
                        c_dp, c_l, c_dp]),
    "sgp_applymap_tok": (c_i, [c_vp, c_i, c_d, c_i, c_i, c_l, c_l, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_l,
                               c_dp, c_dp, c_dp, c_dp, c_l, c_dp, c_dp]),
    "sgp_nll": (c_i, [c_vp, c_i, c_d, c_i, c_dp, c_dp, c_dp, c_l, c_i, c_dp]),
    "sgp_nll_dev": (c_i, [c_vp, c_i, c_d, c_i, c_dp, c_vp, c_vp, c_l, c_i, c_vp]),
    "sgp_fit": (c_i, [c_vp, c_i, c_d, c_i, c_dp, c_dp, c_dp, c_l, c_dp, c_dp, c_dp, c_dp]),
    "sgp_applymap": (c_i, [c_vp, c_i, c_i, c_d, c_i, c_l, c_l, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_l, c_dp,
                           c_dp, c_dp, c_l, c_dp, c_dp, c_dp, c_l, c_dp, c_dp, c_ullp]),
    "sgp_applymap_split": (c_i, [c_vp, c_i, c_d, c_i, c_i, c_l, c_l, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_l, c_dp, c_dp,
                                 c_dp, c_l, c_dp, c_dp, c_ullp]),
    "sgp_model_create": (c_i, [c_vp, c_i, c_d, c_dp, c_dp, c_dp, c_dp, c_dp, c_l, c_dp, c_dp, c_dp, c_l,
                               ctypes.POINTER(c_vp)]),
    "sgp_model_destroy": (c_i, [c_vp]),
    "sgp_model_applymap_dev": (c_i, [c_vp, c_vp, c_i, c_i, c_l, c_l, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_l, c_vp]),
    "sgp_applymap_quality": (c_i, [c_vp, c_i, c_i, c_d, c_i, c_l, c_l, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_l, c_dp,
                                   c_dp, c_dp, c_l, c_i, c_dp, c_l, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ullp]),
    "sgp_applymap4": (c_i, [c_vp, c_l, c_l, c_dp, c_dp, c_dp, c_dp, c_dp, c_l, c_dp, c_dp, c_l, c_dp, c_dp, c_ullp]),
    "sgp_standard_map_iterate": (c_i, [c_vp, c_d, c_l, c_l, c_dp, c_dp]),
    "sgp_model_applymap_quality_dev": (c_i, [c_vp, c_vp, c_i, c_i, c_l, c_l, c_vp, c_vp, c_vp, c_vp, c_i, c_dp, c_l, c_vp,
                                             c_vp, c_vp, c_vp, c_vp, c_vp]),
    "sgp_map_last_passes": (c_i, [c_vp, c_ullp]),
    "sgp_alpha_cache_stats": (c_i, [c_vp, c_ullp, c_ullp]),
    "sgp_compute_r": (c_d, [c_d, c_d, c_d, c_d]),
    "sgp_ath": (c_d, [c_d, c_d, c_d]),
    "sgp_spd_factor": (c_i, [c_vp, c_dp, c_l, c_dp, c_dp, c_dp]),
    "sgp_selftest_gemm": (c_i, [c_vp, c_i, c_i, c_i, c_i, c_i, c_i, c_dp]),
    "sgp_i8mma_selftest": (c_i, [c_vp, c_i, ctypes.POINTER(c_i), ctypes.POINTER(c_i), ctypes.POINTER(c_i)]),
    "sgp_bench_dfma": (c_i, [c_vp, c_i, c_dp]),
    "sgp_set_ozaki": (c_i, [c_vp, c_i]),
    "sgp_set_ozaki_ex": (c_i, [c_vp, c_i, c_i, c_l]),
    "sgp_ozaki_gemm_host_ex": (c_i, [c_vp, c_i, c_l, c_l, c_l, c_d, c_dp, c_l, c_i, c_i, c_dp, c_l, c_i, c_i, c_d, c_dp, c_l, c_i, c_i]),
    "sgp_ozaki_gemm_host": (c_i, [c_vp, c_i, c_l, c_l, c_l, c_d, c_dp, c_l, c_dp, c_l, c_d, c_dp, c_l]),
    "sgp_ozaki_bench": (c_i, [c_vp, c_i, c_l, c_l, c_l, c_i, c_dp]),
    "sgp_gemm_host": (c_i, [c_vp, c_i, c_i, c_i, c_i, c_i, c_i, c_d, c_d, c_dp, c_l, c_dp, c_l, c_dp, c_l]),
    "sgp_bench_gemm": (c_i, [c_vp, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_dp]),
    "sgp_fill_sym_dev": (c_i, [c_vp, c_i, c_d, c_i, c_dp, c_vp, c_l, c_vp, c_l]),
    "sgp_potrf_dev": (c_i, [c_vp, c_vp, c_l, c_l, c_vp]),
    "sgp_build_k_dev": (c_i, [c_vp, c_i, c_d, c_vp, c_vp, c_l, c_vp, c_vp, c_l, c_dp, c_vp, c_l]),
}


def lib():
    """Load (building if necessary) the shared library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        try:
            L = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise RuntimeError(f"cannot load {LIB_PATH}: {e} (sympgpr_b200 has no CPU fallback)") from e
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    msg = lib().sgp_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def device_count():
    return lib().sgp_device_count()


def check(status, what):
    """Map a C status to the exception the reference-side caller would see."""
    if status == 0:
        return
    msg = last_error()
    if status > 0:
        # scipy.linalg.cholesky raises LinAlgError on a non-positive pivot; the scripts' bare
        # `except:` fallbacks (python/02_pert_pendulum/func.py:194-204) depend on it
        raise np.linalg.LinAlgError(f"{what}: {status}-th leading minor of the array is not positive definite")
    if status == E_BADARG:
        raise ValueError(f"{what}: {msg}")
    if status == E_NOMEM:
        raise MemoryError(f"{what}: {msg}")
    if status == E_NODEV:
        raise RuntimeError(f"{what}: no CUDA device -- sympgpr_b200 has no CPU fallback ({msg})")
    raise RuntimeError(f"{what}: CUDA error ({msg})")


class Context:
    """One stream + cached workspaces on one device."""

    def __init__(self, device=0):
        h = c_vp()
        check(lib().sgp_create(int(device), ctypes.byref(h)), "sgp_create")
        self.handle = h
        self.device = int(device)

    def set_stream(self, stream_ptr):
        check(lib().sgp_set_stream(self.handle, c_vp(stream_ptr)), "sgp_set_stream")

    def synchronize(self):
        check(lib().sgp_synchronize(self.handle), "sgp_synchronize")

    def set_ozaki(self, nslices):
        """OPT-IN: 4..8 = the lauum stage of the inverse runs on the INT8 tensor pipe (csrc/ozaki.cu); 0 = DMMA (default)."""
        check(lib().sgp_set_ozaki(self.handle, int(nslices)), "sgp_set_ozaki")

    def set_ozaki_ex(self, nslices, stages=3, leaf_n=0):
        """OPT-IN with the stages named: 1 = lauum, 2 = Cholesky factor + triangular inverse as one recursion of sliced INT8
        products (csrc/ozaki_chol.cu; blocks of <= leaf_n rows stay on the DMMA kernels), 3 = both; nslices = 0 switches off."""
        check(lib().sgp_set_ozaki_ex(self.handle, int(nslices), int(stages), int(leaf_n)), "sgp_set_ozaki_ex")

    def release_workspace(self):
        check(lib().sgp_release_workspace(self.handle), "sgp_release_workspace")

    def close(self):
        if self.handle:
            lib().sgp_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_device():
    return int(os.environ.get("SYMPGPR_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def context(device=None):
    """Process-wide default context of a device (created on first use)."""
    dev = default_device() if device is None else int(device)
    with _lock:
        ctx = _ctx.get(dev)
    if ctx is None:
        ctx = Context(dev)
        oz = os.environ.get("SYMPGPR_B200_OZAKI")
        if oz:
            # "slices[:stages[:leaf_rows]]", e.g. 7:3 -- the opt-in INT8 route for scripts that run unchanged
            f = [int(v) for v in oz.split(":")]
            ctx.set_ozaki_ex(f[0], f[1] if len(f) > 1 else 3, f[2] if len(f) > 2 else 0)
        with _lock:
            _ctx.setdefault(dev, ctx)
            ctx = _ctx[dev]
    return ctx


def dptr(a):
    return a.ctypes.data_as(c_dp)


def as_f64(a):
    """f2py intent(in) coercion: any array-like -> contiguous float64 (copy if needed)."""
    return np.ascontiguousarray(a, dtype=np.float64)


def as_f64_fortran(a):
    return np.asfortranarray(a, dtype=np.float64)
