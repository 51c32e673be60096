"""The C-ABI shared library loads without a GPU and exports every entry point the header declares; the product
package never routes through the oracle."""
import ctypes
import os
import re

from conftest import ROOT

PKG = os.path.join(ROOT, "n_hexane_pyrolysis_surrogate_reactor_model_b200")


def _declared():
    src = open(os.path.join(ROOT, "include", "crnn_pfr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:pfr|crnn)_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from n_hexane_pyrolysis_surrogate_reactor_model_b200 import _lib
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.build import build
    build()
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert _lib.lib().pfr_version() >= 100
    assert _lib.lib().pfr_status_string(-3) == b"workspace too small"


def test_argument_validation_without_a_device():
    """Calls that fail validation return before touching CUDA."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200 import _lib
    L = _lib.lib()
    assert L.pfr_mlp_workspace_bytes(1000, 0) == (4 * 512 + 800) * 1024 * 4
    assert L.pfr_mlp_workspace_bytes(10 ** 6, 0) == (4 * 512 + 800) * (2 * 4 * 148 * 128) * 4   # default chunk: two lanes of 592 row tiles
    assert L.pfr_integrate(None, 0, 64, 4, None, None, None, None, None, None, None, 1e-6, 1e-6, 0, 0, None, None, None, None, None) == -1
    assert L.pfr_integrate(None, 0, 64, 0, None, None, None, None, None, None, None, 1e-6, 1e-6, 0, 0, None, None, None, None, None) == 0
    assert L.pfr_rhs(None, 3, None, None, None, 16, None) == -1
    assert L.pfr_measure_peaks(None) == -1
    # the predictor-MLP trainer and the fast-path integrators validate before touching CUDA as well
    import ctypes
    h = ctypes.c_void_p()
    assert L.pfr_mlp_trainer_create(3, None, None, ctypes.byref(h)) == -1          # in_dim must be 2 or 4
    assert L.pfr_mlp_trainer_create(2, None, None, None) == -1
    assert L.pfr_mlp_trainer_step(None, None, None, 32, 1e-3, 0.9, 0.999, 1e-8, None, None) == -1
    assert L.pfr_mlp_trainer_forward(None, None, 0, None, None) == 0                # empty batch: no-op
    assert L.pfr_mlp_trainer_destroy(None) == 0
    assert L.pfr_integrate(None, 7, 64, 4, None, None, None, None, None, None, None, 1e-6, 1e-6, 0, 0, None, None, None, None, None) == -1
    assert L.pfr_time_grid(None, None, None, None, None, 0, None, None, 0, None, 0, 0, None) == 0     # empty batch
    assert L.pfr_temp_profile(None, None, None, 0, None, 0, None, 0, 0, None) == 0


def test_round2_entry_points_validate_without_a_device():
    """pfr_sweep_*, pfr_stiff_fallback, crnn_model_update, pfr_loss_grad_staged: argument errors are reported before CUDA is
    touched; the size helpers are pure host arithmetic."""
    import ctypes
    import numpy as np
    from n_hexane_pyrolysis_surrogate_reactor_model_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.pfr_sweep_create(None, None, None, 16, ctypes.byref(h)) == -1
    assert L.pfr_sweep_run(None, None, None, None, None, 4, 4, 64, 1e-6, 1e-6, 0, 0, None, None, None, None, None, None, None) == -1
    assert L.pfr_sweep_run(None, None, None, None, None, 0, 4, 64, 1e-6, 1e-6, 0, 0, None, None, None, None, None, None, None) == 0   # empty batch
    assert L.pfr_sweep_destroy(None) == 0
    assert L.pfr_stiff_fallback(None, 3, 64, 4, None, None, None, None, None, None, 1e-6, 1e-6, 0, 0, None, None, None, None, None, None) == -1
    assert L.pfr_reduce_rows_ok(None, 3, 4, None, None, None) == -1
    # one model handle, updated in place (host memory only: works without a GPU)
    w_in, w_b, w_out = np.ones((11, 9), np.float32), np.ones(9, np.float32), np.ones((9, 9), np.float32)
    fp = lambda a: a.ctypes.data_as(_lib.c_float_p)
    assert L.crnn_model_create(fp(w_in), fp(w_b), fp(w_out), None, ctypes.byref(h)) == 0
    assert L.crnn_model_update(h, fp(2 * w_in), fp(w_b), fp(w_out)) == 0
    assert L.crnn_model_update(h, None, fp(w_b), fp(w_out)) == -1
    assert L.crnn_model_destroy(h) == 0
    # workspace arithmetic: 3201 nodes x 130 doubles + 6400 stages x 9 doubles per condition at two sub-steps
    assert L.pfr_loss_grad_workspace_bytes(640, 2) == (3201 * 130 + 6400 * 9) * 640 * 8
    assert L.pfr_loss_grad_workspace_bytes(0, 2) == 0
    assert L.pfr_loss_grad_staged(None, 4, None, None, None, None, None, None, 2, None, None, None, 0, None) == -1
    eon = L.pfr_sweep_device_bytes(1 << 20, 1)
    assert 11.0e9 < eon < 12.5e9 and L.pfr_sweep_device_bytes(1 << 20, 0) < eon / 5    # two [801][n] grids (6.7 GB) + three MLP workspaces (1.7 GB each)


def test_missing_library_fails_loudly(monkeypatch):
    import pytest
    from n_hexane_pyrolysis_surrogate_reactor_model_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("CRNN_PFR_LIB", "/nonexistent/libcrnn.so")
    with pytest.raises(_lib.PfrError):
        _lib.lib()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU", ""), os.path.join(dirpath, f)
