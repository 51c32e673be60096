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
