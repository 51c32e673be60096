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


def test_sass_of_the_built_library_uses_the_hardware_paths_the_design_claims():
    """Static evidence from the shipped binary (cuobjdump, no GPU needed).  Two properties of the build are decided by ptxas, not by
    the source, so they are pinned here: (1) the explicit integrators feed their CRNN coefficients through UNIFORM registers (LDCU.64
    + DFMA R, R, UR, R) -- ptxas falls back to per-thread constant loads (LDC.64 with a vector-register index) as soon as it loses
    track of the warp being converged or of the copy index being uniform, which costs 15 % of the kernel; (2) the predictor MLPs run
    on tcgen05 (UTCHMMA) with TMA loads and TMEM read-back, in both operand formats."""
    import shutil
    import subprocess
    import pytest
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.build import LIB, build
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    build()
    sass = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per[name] = {}
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            per[name][op] = per[name].get(op, 0) + 1

    def count(kernel_substrings, op_prefix):
        ks = [k for k in per if all(sub in k for sub in kernel_substrings)]
        assert ks, kernel_substrings
        return min(sum(v for o, v in per[k].items() if o.startswith(op_prefix)) for k in ks)

    for kern in (["bs23_kernelId"], ["dp54_kernelId"]):
        assert count(kern, "LDCU.64") >= 162, (kern, "coefficient loads are not uniform-register loads")
        ks = [k for k in per if kern[0] in k]
        assert max(per[k].get("LDC.64", 0) for k in ks) < 40, (kern, "per-thread constant loads")
    for fmt in ("ILb0ELb0E", "ILb0ELb1E", "ILb1ELb0E", "ILb1ELb1E"):   # <kFinal, kHalf>
        kern = ["mlp_tc_gemm_kernel", fmt]
        assert count(kern, "UTCHMMA") >= 12 and count(kern, "UTMALDG") >= 4 and count(kern, "LDTM") >= 8
