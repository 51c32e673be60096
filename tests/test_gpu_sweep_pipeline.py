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


def test_free_stepping_dp54_warp_dense_output(surrogates, golden, conditions):
    """PFR_METHOD_DP54_WARP (the isothermal training step's forward pass): free Dormand-Prince steps over tgrid's span with the 801
    knot states read off the method's 4th-order continuous extension.  Against the converged oracle solution (golden fixtures:
    outlet and every 50th knot) within 1e-6 at rtol 1e-8 / atol 1e-11; against the knot-limited BS23 pass at a tight tolerance on
    the 400 shipped conditions within 1e-6 at EVERY knot (1e-3 mol/m3 floor) at 1e-9 / 1e-12 and within 1e-5 at the trainer's
    setting; a few dozen steps instead of 800; outlet knots
    short of the grid end (idx_end), degenerate ones (idx_end = 0) and ragged batch sizes handled as the other integrators do;
    a temperature profile is refused (the method is isothermal)."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200 import _lib
    from conftest import rel_err
    s = surrogates("LLNL", "Eoff")
    tgd = torch.as_tensor(golden["Eoff/tgrid"].T.copy()).cuda()
    r = s.integrate(golden["T"], golden["c0"][:, 6], method="dp54w", tgrid=tgd, rtol=1e-8, atol=1e-11, dense=True, stiff_fallback=None)
    assert int(r.status.abs().sum()) == 0
    truth = np.clip(golden["Eoff/truth_outlet"], 1e-6, 60.0)
    assert np.max(rel_err(r.y.cpu().numpy().T, truth)) < 1e-6
    dense = r.dense.cpu().numpy()
    for j, k in enumerate(range(0, 801, 50)):
        ref = np.clip(golden["Eoff/truth_knots_every50"][:, j, :], 1e-6, 60.0)
        assert np.max(rel_err(dense[k].T, ref)) < 1e-6, k
    st = r.stats.cpu().numpy()
    assert np.array_equal(st[2], 6 * (st[0] + st[1]) + 1)
    assert (st[0] + st[1]).max() < 400

    T, P, L, U = cond4(conditions)
    ref = s.sweep(T, P, L, U, method="rodas4", keep_grids=True)      # for the grids
    c0 = s.inlet_concentration(T, P)
    idx = torch.full((len(T),), 800, dtype=torch.int32, device="cuda")
    idx[::7] = 0
    idx[1::7] = 333
    kw = dict(tgrid=ref.tgrid, idx_end=idx, dense=True, dense_raw=True)
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import FREE_STEP_TOLERANCE
    a = s.integrate(T, c0, method="bs23", rtol=1e-10, atol=1e-13, **kw)
    scale = a.dense.abs().clamp(min=1e-3)
    c = s.integrate(T, c0, method="dp54w", rtol=1e-9, atol=1e-12, **kw)
    assert int(a.status.abs().sum()) == 0 and int(c.status.abs().sum()) == 0
    assert float(((a.dense - c.dense).abs() / scale).max()) < 1e-6
    # at the trainer's setting (the reference trains at 1e-4 / 1e-6): every knot state within 1e-5 (measured 3.7e-6), outlets within 1e-6
    b = s.integrate(T, c0, method="dp54w", rtol=FREE_STEP_TOLERANCE[0], atol=FREE_STEP_TOLERANCE[1], **kw)
    assert int(b.status.abs().sum()) == 0
    assert float(((a.dense - b.dense).abs() / scale).max()) < 1e-5
    assert float(((a.y - b.y).abs() / a.y.abs().clamp(min=1e-3)).max()) < 1e-5
    attempts = (b.stats[0] + b.stats[1]).float()
    assert float(attempts[idx == 800].mean()) < 120
    print(f"dp54w: {float(attempts[idx == 800].mean()):.0f} attempts per trajectory (bs23: {float((a.stats[0] + a.stats[1]).float()[idx == 800].mean()):.0f})")
    for m in (1, 3, 5):
        sub = s.integrate(T[:m], c0[:m], method="dp54w", tgrid=ref.tgrid[:, :m].contiguous(), idx_end=idx[:m].contiguous(), rtol=1e-7, atol=1e-10)
        assert torch.equal(sub.y, b.y[:, :m])
    with pytest.raises(_lib.PfrError):
        s.integrate(T, c0, method="dp54w", tgrid=ref.tgrid, Tprof=ref.tgrid, rtol=1e-7, atol=1e-10)
