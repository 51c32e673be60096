"""Pins for row a8: the oracle's restatement of torchdiffeq's dopri5.

torchdiffeq itself is not installable in the build container (no network), so the strongest pin -- the real library run
on the golden conditions -- is written to engage by itself the moment it can:

  * test_oracle_matches_live_torchdiffeq         pytest.importorskip("torchdiffeq"): oracle vs the library, 16 golden conditions
  * test_oracle_matches_stored_torchdiffeq_vectors   tests/golden/torchdiffeq_vectors.npz, if a reference maintainer has
                                                     produced it with tests/golden/make_torchdiffeq_golden.py

What CAN be pinned here against an independent third-party implementation is pinned against scipy's RK45, which is the
same Dormand-Prince 5(4) pair with the same Hairer starting step (scipy/integrate/_ivp/rk.py, common.py): the tableau, one
Runge-Kutta step with its embedded error, and the initial step size.  (The step-size controller and the dense output
differ between the two libraries by design -- scipy may shrink an accepted step and clips the last one, torchdiffeq does
neither -- so whole trajectories agree only to the tolerance, which is checked as well.)
"""
import os

import numpy as np
import pytest
import torch

from oracle import reference_path as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _golden(variant):
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
    vec = np.load(os.path.join(GOLD, "reference_vectors.npz"))
    ms = ModelSet.from_packed(os.path.join(GOLD, "containers", "LLNL.npz"), variant)
    tg = vec["Eoff/tgrid"] if variant == "Eoff" else vec["Eon/tgrid_full"]
    Tp = np.repeat(vec["T"][:, None], R.NTOTAL, 1) if variant == "Eoff" else vec["Eon/Tprof"]
    return vec, ms.crnn, tg, Tp


def _oracle_run(variant, dtype):
    vec, cr, tg, Tp = _golden(variant)
    sols, nfe = [], []
    for i in range(tg.shape[0]):
        st = R.SolveStats()
        sols.append(R.crnn_predict(tg[i], Tp[i], vec["c0"][i], cr.w_in, cr.w_b, cr.w_out, dtype=dtype, stats=st))
        nfe.append(st.nfe)
    return np.stack(sols), np.array(nfe)


# ------------------------------------------------------------------------------------------------ the real library
@pytest.mark.parametrize("variant", ["Eoff", "Eon"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_matches_live_torchdiffeq(variant, dtype):
    """Reference call sites: SURROGATE_MODEL/surrogate_model_Eoff_single_model.py:185, ...Eon_single_model.py:154-155."""
    td = pytest.importorskip("torchdiffeq")
    torch.set_num_threads(1)
    vec, cr, tg, Tp = _golden(variant)
    mine, nfe = _oracle_run(variant, dtype)
    for i in range(tg.shape[0]):
        t = torch.tensor(tg[i], dtype=dtype)
        calls = [0]
        f0 = R.CRNNFunc(t, torch.tensor(Tp[i], dtype=dtype), *(torch.tensor(a, dtype=dtype) for a in (cr.w_in, cr.w_b, cr.w_out)))

        def f(tt, y):
            calls[0] += 1
            return f0(tt, y)
        with torch.no_grad():
            sol = td.odeint(f, torch.tensor(vec["c0"][i], dtype=dtype), t, method="dopri5", atol=1e-6, rtol=1e-6)
        ref = torch.clamp(sol.T, R.LB, R.UB).numpy()
        assert calls[0] == nfe[i], (variant, i, calls[0], nfe[i])       # same step sequence
        np.testing.assert_allclose(mine[i], ref, rtol=1e-6, atol=1e-9)  # same arithmetic (bit-level in practice)


@pytest.mark.parametrize("variant", ["Eoff", "Eon"])
def test_oracle_matches_stored_torchdiffeq_vectors(variant):
    path = os.path.join(GOLD, "torchdiffeq_vectors.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/torchdiffeq_vectors.npz not produced yet (run tests/golden/make_torchdiffeq_golden.py where "
                    "torchdiffeq is installed): a8 stays PARITY UNPINNED against the real library")
    gold = np.load(path)
    for name, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        mine, nfe = _oracle_run(variant, dtype)
        assert np.array_equal(nfe, gold[f"{variant}/nfe_{name}"]), (variant, name)
        np.testing.assert_allclose(mine, gold[f"{variant}/dopri5_{name}"], rtol=1e-6, atol=1e-9)


# ------------------------------------------------------------------------------------------------ scipy's RK45
def test_tableau_equals_scipy_rk45():
    from scipy.integrate import RK45
    alpha, beta, c_sol, c_err = R.dopri5_tableau(torch.float64)
    assert np.allclose([float(a) for a in alpha], list(RK45.C[1:]) + [1.0], rtol=0, atol=1e-16)
    for s, row in enumerate(beta[:5]):
        assert np.allclose(row.numpy(), RK45.A[s + 1][: s + 1], rtol=0, atol=1e-16), s
    assert np.allclose(beta[5].numpy(), RK45.B, rtol=0, atol=1e-16)          # 7th stage row = the 5th-order weights (FSAL)
    assert np.allclose(c_sol.numpy()[:6], RK45.B, rtol=0, atol=1e-16)
    # Embedded error weights: both libraries use a multiple of the Dormand-Prince vector b - b^ = (71/57600, 0, -71/16695,
    # 71/1920, -17253/339200, 22/525, -1/40).  scipy stores -(b - b^); torchdiffeq's dopri5.py stores the variant built on
    # (1951/21600, 0, 22642/50085, 451/720, -12231/42400, 649/6300, 1/60), which is exactly 2/3 of it -- i.e. torchdiffeq's
    # error ratio is 2/3 of scipy's for the same step, a documented property of the reference's solver that the oracle keeps.
    assert np.allclose(-1.5 * c_err.numpy(), RK45.E, rtol=0, atol=1e-16)


@pytest.mark.parametrize("variant", ["Eoff", "Eon"])
def test_rk_step_and_initial_step_equal_scipy(variant):
    """One dopri5 step (solution, FSAL slope, embedded error) and the Hairer starting step against scipy's rk_step /
    select_initial_step, float64, on the golden conditions at the inlet state and at a mid-trajectory state."""
    from scipy.integrate import RK45
    from scipy.integrate._ivp.common import select_initial_step
    from scipy.integrate._ivp.rk import rk_step as scipy_rk_step
    vec, cr, tg, Tp = _golden(variant)
    alpha, beta, _, c_err = R.dopri5_tableau(torch.float64)
    for i in range(0, tg.shape[0], 3):
        t = torch.tensor(tg[i], dtype=torch.float64)
        func = R.CRNNFunc(t, torch.tensor(Tp[i], dtype=torch.float64), *(torch.tensor(a, dtype=torch.float64) for a in (cr.w_in, cr.w_b, cr.w_out)))
        f = lambda tt, y, prev=False: func(tt, y)                      # (stage times exactly as handed over: no nextafter)
        fnp = lambda tt, y: func(torch.tensor(tt, dtype=torch.float64), torch.tensor(y)).numpy()
        y0 = torch.tensor(vec["c0"][i], dtype=torch.float64)
        f0 = f(t[0], y0)
        dt = R.select_initial_step(f, t[0], y0, f0, 1e-6, 1e-6)
        h_scipy = select_initial_step(fnp, 0.0, y0.numpy(), np.inf, np.inf, f0.numpy(), 1, RK45.error_estimator_order, 1e-6, 1e-6)
        assert abs(float(dt) - h_scipy) <= 1e-12 * h_scipy, (i, float(dt), h_scipy)
        mid = torch.tensor(vec[f"{variant}/truth_knots_every50"][i, 4], dtype=torch.float64)   # the state at knot 200
        for ya, ta, h in ((y0, 0.0, float(dt)), (mid, float(t[200]), 0.37 * float(t[201] - t[200]))):
            ta_t, h_t = torch.tensor(ta, dtype=torch.float64), torch.tensor(h, dtype=torch.float64)
            fa = f(ta_t, ya)
            yb, fb, err, _ = R.rk_step(f, ya, fa, ta_t, h_t, alpha, beta, c_err)
            K = np.empty((7, 9))
            y_new, f_new = scipy_rk_step(fnp, ta, ya.numpy(), fa.numpy(), h, RK45.A, RK45.B, RK45.C, K)
            scale = np.maximum(np.abs(y_new), 1e-3)
            assert np.max(np.abs(yb.numpy() - y_new) / scale) < 1e-13
            assert np.max(np.abs(fb.numpy() - f_new) / np.maximum(np.abs(f_new), 1e-3)) < 1e-11
            e_scipy = K.T @ RK45.E * h * (-2.0 / 3.0)   # see test_tableau_equals_scipy_rk45
            # the error is a difference of nearly equal sums: compare on the scale of its terms (h |K| |E|), not of the result
            assert np.max(np.abs(err.numpy() - e_scipy)) <= 1e-13 * h * np.max(np.abs(K)) * np.sum(np.abs(RK45.E))


def test_isothermal_trajectory_agrees_with_scipy_rk45_to_the_tolerance():
    """Whole-trajectory sanity against an independent adaptive DP5(4): same pair, different controller, so agreement is at the
    level the common tolerance (1e-6) allows, not bitwise."""
    from scipy.integrate import solve_ivp
    vec, cr, tg, Tp = _golden("Eoff")
    worst = 0.0
    for i in (0, 5, 11):
        mine = R.crnn_predict(tg[i], Tp[i], vec["c0"][i], cr.w_in, cr.w_b, cr.w_out, dtype=torch.float64)[:, -1]
        rhs = lambda tt, y, i=i: R.crnn_rhs_np(np.array([Tp[i, 0]], np.float64), y[None, :], cr.w_in, cr.w_b, cr.w_out)[0]
        s = solve_ivp(rhs, (0.0, float(tg[i, -1])), vec["c0"][i].astype(np.float64), method="RK45", rtol=1e-6, atol=1e-6)
        assert s.success
        worst = max(worst, float(np.max(np.abs(np.clip(s.y[:, -1], R.LB, R.UB) - mine) / np.maximum(np.abs(mine), 1e-3))))
    assert worst < 1e-3, worst   # each is up to 8e-4 from the converged solution at this tolerance
