"""ctypes binding of the C-ABI library (include/crnn_pfr.h).  There is no CPU fallback: if the CUDA
extension is missing or a call fails, this module raises."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcrnn_pfr_b200.so")

PFR_OK = 0
METHOD_RODAS4 = 0
METHOD_DOPRI5 = 1
METHOD_RODAS4_TPC = 2
METHOD_ROS3 = 3
METHOD_BS23 = 4
METHOD_DP54 = 5
METHOD_BS23_WARP = 6
METHOD_DP54_WARP = 8
METHOD_TAYLOR4 = 7
ST_STIFF = 4
SWEEP_NO_ORDER, SWEEP_NO_FALLBACK = 1, 2
STATUS_TEXT = {0: "ok", 1: "max steps exceeded", 2: "non-finite state", 3: "step size underflow",
               4: "stiff for the explicit fast path (integrate with ros3 / rodas4)"}

c_void_p, c_int, c_double, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t
c_float_p = ctypes.POINTER(ctypes.c_float)
c_double_p = ctypes.POINTER(ctypes.c_double)

_lib = None


class PfrError(RuntimeError):
    pass


def _declare(lib):
    lib.pfr_version.restype = c_int
    lib.pfr_status_string.restype = ctypes.c_char_p
    lib.pfr_status_string.argtypes = [c_int]
    lib.pfr_last_cuda_error.restype = ctypes.c_char_p
    lib.pfr_launch_count.restype = ctypes.c_ulonglong
    lib.crnn_model_create.argtypes = [c_float_p, c_float_p, c_float_p, c_double_p, ctypes.POINTER(c_void_p)]
    lib.crnn_model_destroy.argtypes = [c_void_p]
    lib.crnn_model_update.argtypes = [c_void_p, c_float_p, c_float_p, c_float_p]
    lib.crnn_model_update.restype = c_int
    lib.pfr_sweep_create.argtypes = [c_void_p, c_void_p, c_void_p, c_int, ctypes.POINTER(c_void_p)]
    lib.pfr_sweep_destroy.argtypes = [c_void_p]
    lib.pfr_sweep_device_bytes.argtypes = [c_int, c_int]
    lib.pfr_sweep_device_bytes.restype = c_size_t
    lib.pfr_sweep_run.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.pfr_sweep_stiff_count.argtypes = [c_void_p, ctypes.POINTER(c_int)]
    lib.pfr_sweep_integrator_ms.argtypes = [c_void_p, ctypes.POINTER(ctypes.c_float)]
    for name in ("pfr_sweep_create", "pfr_sweep_destroy", "pfr_sweep_run", "pfr_sweep_stiff_count", "pfr_sweep_integrator_ms"):
        getattr(lib, name).restype = c_int
    lib.pfr_mlp_create.argtypes = [c_int, ctypes.POINTER(c_float_p), ctypes.POINTER(c_float_p), c_double, c_double,
                                   c_double_p, c_double_p, ctypes.POINTER(c_void_p)]
    lib.pfr_mlp_destroy.argtypes = [c_void_p]
    lib.pfr_mlp_set_mode.argtypes = [c_void_p, c_int]
    lib.pfr_mlp_set_mode.restype = c_int
    lib.pfr_mlp_workspace_bytes.restype = c_size_t
    lib.pfr_mlp_workspace_bytes.argtypes = [c_int, c_int]
    lib.pfr_inlet_concentration.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    lib.pfr_time_grid.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                  c_void_p, c_size_t, c_int, c_void_p]
    lib.pfr_temp_profile.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_size_t, c_int,
                                     c_void_p]
    lib.pfr_idx_cut.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    lib.pfr_rhs.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    lib.pfr_integrate.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_double, c_double, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p]
    lib.pfr_stiff_fallback.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double,
                                       c_double, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.pfr_stiff_fallback.restype = c_int
    lib.pfr_loss_grad.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                  c_void_p, c_void_p]
    lib.pfr_loss_grad.restype = c_int
    lib.pfr_loss_grad_workspace_bytes.argtypes = [c_int, c_int]
    lib.pfr_loss_grad_workspace_bytes.restype = c_size_t
    lib.pfr_loss_grad_staged.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                         c_void_p, c_void_p, c_size_t, c_void_p]
    lib.pfr_loss_grad_staged.restype = c_int
    lib.pfr_reduce_rows.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p]
    lib.pfr_reduce_rows.restype = c_int
    lib.pfr_reduce_rows_ok.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.pfr_reduce_rows_ok.restype = c_int
    lib.pfr_accuracy.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]
    lib.pfr_accuracy.restype = c_int
    fpp = ctypes.POINTER(c_float_p)
    lib.pfr_mlp_trainer_create.argtypes = [c_int, fpp, fpp, ctypes.POINTER(c_void_p)]
    lib.pfr_mlp_trainer_destroy.argtypes = [c_void_p]
    lib.pfr_mlp_trainer_step.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_double, c_double, c_double, c_double, c_void_p, c_void_p]
    lib.pfr_mlp_trainer_forward.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    lib.pfr_mlp_trainer_loss.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    lib.pfr_mlp_trainer_read.argtypes = [c_void_p, fpp, fpp]
    for name in ("pfr_mlp_trainer_create", "pfr_mlp_trainer_destroy", "pfr_mlp_trainer_step", "pfr_mlp_trainer_forward",
                 "pfr_mlp_trainer_loss", "pfr_mlp_trainer_read"):
        getattr(lib, name).restype = c_int
    lib.pfr_measure_peaks.argtypes = [c_double_p]
    lib.pfr_fastmath.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.pfr_fastmath.restype = c_int
    for name in ("crnn_model_create", "crnn_model_destroy", "pfr_mlp_create", "pfr_mlp_destroy",
                 "pfr_inlet_concentration", "pfr_time_grid", "pfr_temp_profile", "pfr_idx_cut", "pfr_rhs",
                 "pfr_integrate", "pfr_measure_peaks"):
        getattr(lib, name).restype = c_int


def lib():
    """The loaded shared library.  Raises if it has not been built (python -m <package>.build)."""
    global _lib
    if _lib is None:
        path = os.environ.get("CRNN_PFR_LIB", LIB_PATH)
        if not os.path.exists(path):
            raise PfrError(f"{path} is missing: build the CUDA extension first "
                           f"(python -m n_hexane_pyrolysis_surrogate_reactor_model_b200.build); there is no CPU fallback")
        _lib = ctypes.CDLL(path)
        _declare(_lib)
    return _lib


def check(rc: int, what: str):
    if rc != PFR_OK:
        L = lib()
        msg = L.pfr_status_string(rc).decode()
        if rc == -2:
            msg += ": " + L.pfr_last_cuda_error().decode()
        raise PfrError(f"{what} failed ({rc}): {msg}")
