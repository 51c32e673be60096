import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "reference_vectors.npz"))


@pytest.fixture(scope="session")
def conditions():
    return np.load(os.path.join(GOLDEN, "conditions.npz"))


def packed_path(mech):
    return os.path.join(GOLDEN, "containers", f"{mech}.npz")


@pytest.fixture(scope="session")
def model_sets():
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet

    cache = {}

    def get(mech="LLNL", variant="Eoff", crnn_key=None):
        key = (mech, variant, crnn_key)
        if key not in cache:
            cache[key] = ModelSet.from_packed(packed_path(mech), variant, crnn_key)
        return cache[key]

    return get


@pytest.fixture(scope="session")
def surrogates(model_sets):
    """GPU-resident Surrogate objects, created on first use."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate

    cache = {}

    def get(mech="LLNL", variant="Eoff", crnn_key=None, mlp_mode="f16x3"):
        key = (mech, variant, crnn_key, mlp_mode)
        if key not in cache:
            cache[key] = Surrogate(model_sets(mech, variant, crnn_key), mlp_mode=mlp_mode)
        return cache[key]

    return get


def cond4(conditions, name="independent_4D", n=None):
    a = conditions[name]
    if n is not None:
        a = a[:n]
    T = a[:, 0].astype(np.float32)
    P = (a[:, 1] * 1e5).astype(np.float32)
    if a.shape[1] == 4:
        L, U = a[:, 2].astype(np.float32), a[:, 3].astype(np.float32)
    else:
        L, U = np.full_like(T, 1.0), np.full_like(T, 2.5)
    return T, P, L, U


def rel_err(a, b, floor=1e-3):
    """Relative error with the survey's species floor (1e-3 mol/m3)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
