"""The one-call sweep (pfr_sweep_run) against the stage-by-stage path it replaces: same kernels, same arithmetic per condition,
so the results must be IDENTICAL bit for bit whatever order the conditions are visited in."""
import dataclasses

import numpy as np
import pytest
import torch

from conftest import cond4

pytestmark = pytest.mark.gpu


def _lhs(n, seed=7):
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
    return lhs_conditions(n, seed=seed)


@pytest.mark.parametrize("variant,method,tol", [("Eon", "bs23", (3e-7, 1e-12)), ("Eon", "taylor4", (3e-7, 1e-12)), ("Eon", "ros3", (1e-7, 1e-7)), ("Eon", "rodas4", (1e-6, 1e-6)),
                                                ("Eoff", "dp54", (1e-7, 1e-7)), ("Eoff", "rodas4", (1e-6, 1e-6))])
@pytest.mark.parametrize("n", [1, 257, 5000])
def test_one_call_sweep_equals_staged_sweep(surrogates, variant, method, tol, n):
    T, P, L, U = _lhs(n)
    s = surrogates("LLNL", variant)
    kw = dict(method=method, rtol=tol[0], atol=tol[1])
    one = s.sweep(T, P, L, U, staged=False, **kw)
    ref = s.sweep(T, P, L, U, staged=True, **kw)
    assert int((one.status != 0).sum()) == 0
    assert torch.equal(one.y, ref.y) and torch.equal(one.status, ref.status) and torch.equal(one.stats, ref.stats)
    assert torch.equal(one.t_end, ref.t_end)
    if variant == "Eon":
        assert torch.equal(one.idx_cut, ref.idx_cut)
    assert one.stiff_fallbacks == 0


def test_one_call_sweep_full_length_grid_and_float32(surrogates, conditions):
    """sampling_case_2D.csv style input (no L / u0: outlet at the last knot of the full-length grid) and float32 state."""
    a = conditions["independent_2D"]
    T, P = a[:, 0].astype(np.float32), (a[:, 1] * 1e5).astype(np.float32)
    s = surrogates("LLNL", "Eon")
    one = s.sweep(T, P, method="bs23", rtol=1e-8, atol=1e-8, staged=False)
    ref = s.sweep(T, P, method="bs23", rtol=1e-8, atol=1e-8, staged=True)
    assert torch.equal(one.y, ref.y) and int((one.idx_cut != 800).sum()) == 0
    T, P, L, U = cond4(conditions)
    for variant, method in (("Eon", "bs23"), ("Eoff", "dp54")):
        s = surrogates("LLNL", variant)
        one = s.sweep(T, P, L, U, method=method, precision=32, rtol=1e-6, atol=1e-6, staged=False)
        ref = s.sweep(T, P, L, U, method=method, precision=32, rtol=1e-6, atol=1e-6, staged=True)
        assert one.y.dtype == torch.float32 and torch.equal(one.y, ref.y)


def test_one_call_sweep_stiff_fallback_stays_on_the_device(model_sets):
    """Rates 3000x faster than trained (ln A + 8) make the knot intervals stiff for the explicit method: the fast path flags those
    conditions, the pipeline's device-side list hands exactly them to the Rosenbrock kernel, and the result equals the staged
    path's (which finds them with a host-side check)."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200 import _lib
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
    ms = model_sets("LLNL", "Eon")
    fast_crnn = dataclasses.replace(ms.crnn, w_b=(ms.crnn.w_b + 8.0).astype(np.float32))
    s = Surrogate(dataclasses.replace(ms, crnn=fast_crnn))
    T, P, L, U = _lhs(3000, seed=3)
    bare = s.sweep(T, P, L, U, method="bs23", rtol=1e-6, atol=1e-6, staged=True, keep_grids=True)   # staged path, to see the flags
    # (the staged path with its fallback on is the comparison; the flags come from a run of the bare kernel)
    c0 = s.inlet_concentration(T, P)
    flagged = s.integrate(T, c0, tgrid=bare.tgrid, Tprof=bare.Tprof, idx_end=bare.idx_cut, method="bs23", rtol=1e-6, atol=1e-6,
                          stiff_fallback=None).status == _lib.ST_STIFF
    assert int(flagged.sum()) > 0
    one = s.sweep(T, P, L, U, method="bs23", rtol=1e-6, atol=1e-6, staged=False)
    assert one.stiff_fallbacks == int(flagged.sum()) == bare.stiff_fallbacks
    assert int((one.status != 0).sum()) == 0
    assert torch.equal(one.y, bare.y)


def test_one_call_sweep_is_repeatable_and_reuses_its_buffers(surrogates):
    T, P, L, U = _lhs(20000, seed=11)
    s = surrogates("LLNL", "Eon")
    a = s.sweep(T, P, L, U, method="fast", staged=False)
    before = torch.cuda.memory_allocated()
    b = s.sweep(T, P, L, U, method="fast", staged=False)
    assert torch.equal(a.y, b.y)
    assert torch.cuda.memory_allocated() - before < 4 * 20000 * 9 * 8   # only the result tensors are new: the grids live in the handle
    assert s.integrator_ms() > 0.0


def test_handles_refuse_a_foreign_device(surrogates):
    """Library state is per device; a handle used with another device current is refused (needs two GPUs)."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200 import _lib
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible")
    s = surrogates("LLNL", "Eon")
    T, P, L, U = _lhs(64)
    with torch.cuda.device(1):
        with pytest.raises(_lib.PfrError):
            s.time_grid(torch.as_tensor(T, device="cuda:1"), torch.as_tensor(P, device="cuda:1"))


@pytest.mark.parametrize("variant", ["Eoff", "Eon"])
def test_warp_per_condition_bs23_equals_thread_per_condition(surrogates, golden, conditions, variant):
    """PFR_METHOD_BS23_WARP (lane = species, the training step's forward pass) against PFR_METHOD_BS23: the same method and
    controller with the dot products summed in another order -- identical step sequences, dense knot states and outlets equal to
    1e-11 of the state's scale; degenerate outlet knots (idx_end = 0) and ragged batch sizes included."""
    s = surrogates("LLNL", variant)
    T, P, L, U = cond4(conditions)
    ref = s.sweep(T, P, L, U, method="rodas4", keep_grids=True)      # for the grids
    c0 = s.inlet_concentration(T, P)
    idx = ref.idx_cut.clone() if variant == "Eon" else torch.full((len(T),), 800, dtype=torch.int32, device="cuda")
    idx[::7] = 0
    kw = dict(tgrid=ref.tgrid, Tprof=ref.Tprof, idx_end=idx, rtol=1e-8, atol=1e-10, dense=True, dense_raw=True)
    a = s.integrate(T, c0, method="bs23", **kw)
    b = s.integrate(T, c0, method="bs23w", **kw)
    assert int(a.status.abs().sum()) == 0 and int(b.status.abs().sum()) == 0
    assert torch.equal(a.stats[:2], b.stats[:2])                     # accepted / rejected steps
    assert torch.equal(b.stats[2], 3 * (b.stats[0] + b.stats[1]) + (idx != 0).int())
    scale = a.dense.abs().amax(dim=(0, 2), keepdim=True).clamp(min=1e-3)
    assert float(((a.dense - b.dense).abs() / scale).max()) < 1e-11
    assert float(((a.y - b.y).abs() / a.y.abs().clamp(min=1e-3)).max()) < 1e-11
    for m in (1, 3, 5):
        sub = s.integrate(T[:m], c0[:m], method="bs23w", tgrid=ref.tgrid[:, :m].contiguous(), Tprof=None if ref.Tprof is None else ref.Tprof[:, :m].contiguous(),
                          idx_end=idx[:m].contiguous(), rtol=1e-8, atol=1e-10)
        assert torch.equal(sub.y, b.y[:, :m])
