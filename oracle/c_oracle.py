"""CPU ORACLE (test infrastructure, NOT product code): ctypes front-end of oracle/liboracle.so
(crnn_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("crnn_oracle.c", "dopri5_impl.inc")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


class use_fma_build:
    """Context manager: route the calls below through the FMA-contracted build (ill-conditioning demo only)."""

    def __enter__(self):
        global _LIB
        so = os.path.join(_HERE, "liboracle_fma.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "liboracle_fma.so"], stdout=subprocess.DEVNULL)
        self._saved = lib()
        _LIB = ctypes.CDLL(so)
        return self

    def __exit__(self, *exc):
        global _LIB
        _LIB = self._saved
        return False


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def _prep(tgrid, Tprof, u0, w_in, w_b, w_out):
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    tgrid, Tprof, u0 = f(tgrid), f(Tprof), f(u0)
    assert tgrid.ndim == 2 and tgrid.shape[1] == 801 and Tprof.shape == tgrid.shape and u0.shape == (tgrid.shape[0], 9)
    return tgrid, Tprof, u0, f(w_in), f(w_b), f(w_out)


def dopri5_batch(tgrid, Tprof, u0, w_in, w_b, w_out, rtol=1e-6, atol=1e-6, inter=(-30.0, 30.0), precision=32,
                 report=None, dense=False, nthreads=0):
    """Reference-behaviour solve. Returns (y_out[N,9] f64 clamped, sol[N,801,9] f64 | None, stats[N,4] i32)."""
    tgrid, Tprof, u0, w_in, w_b, w_out = _prep(tgrid, Tprof, u0, w_in, w_b, w_out)
    N = tgrid.shape[0]
    y = np.zeros((N, 9), np.float64)
    sol = np.zeros((N, 801, 9), np.float64) if dense else None
    stats = np.zeros((N, 4), np.int32)
    rep = None if report is None else np.ascontiguousarray(report, dtype=np.int32)
    F, D, I = ctypes.c_float, ctypes.c_double, ctypes.c_int
    lib().oracle_dopri5_batch(I(N), _p(tgrid, F), _p(Tprof, F), _p(u0, F), _p(w_in, F), _p(w_b, F), _p(w_out, F),
                              D(rtol), D(atol), D(inter[0]), D(inter[1]), I(precision), _p(rep, I), _p(y, D),
                              _p(sol, D), _p(stats, I), I(nthreads))
    return y, sol, stats


def truth_batch(tgrid, Tprof, u0, w_in, w_b, w_out, inter=(-30.0, 30.0), upto=None, rtol=1e-13, atol=1e-16,
                knots=False, nthreads=0):
    """Converged float64 solution. Returns (y_out[N,9] unclamped, y_knots[N,801,9] | None)."""
    tgrid, Tprof, u0, w_in, w_b, w_out = _prep(tgrid, Tprof, u0, w_in, w_b, w_out)
    N = tgrid.shape[0]
    y = np.zeros((N, 9), np.float64)
    yk = np.zeros((N, 801, 9), np.float64) if knots else None
    up = None if upto is None else np.ascontiguousarray(upto, dtype=np.int32)
    F, D, I = ctypes.c_float, ctypes.c_double, ctypes.c_int
    bad = lib().oracle_truth_batch(I(N), _p(tgrid, F), _p(Tprof, F), _p(u0, F), _p(w_in, F), _p(w_b, F), _p(w_out, F),
                                   D(inter[0]), D(inter[1]), _p(up, I), D(rtol), D(atol), _p(y, D), _p(yk, D), I(nthreads))
    if bad:
        raise RuntimeError(f"oracle_truth_batch: {bad} interval(s) failed")
    return y, yk


def truth_knots_dp(tgrid, Tprof, u0, w_in, w_b, w_out, inter=(-30.0, 30.0), rtol=1e-13, atol=1e-16, nthreads=0):
    """Converged knot states [N,801,9] for DOUBLE parameters (finite differences of the training loss)."""
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    tgrid, Tprof, u0 = f(tgrid), f(Tprof), f(u0)
    w_in, w_b, w_out = d(w_in), d(w_b), d(w_out)
    N = tgrid.shape[0]
    y = np.zeros((N, 9), np.float64)
    yk = np.zeros((N, 801, 9), np.float64)
    F, D, I = ctypes.c_float, ctypes.c_double, ctypes.c_int
    bad = lib().oracle_truth_batch_dp(I(N), _p(tgrid, F), _p(Tprof, F), _p(u0, F), _p(w_in, D), _p(w_b, D), _p(w_out, D),
                                      D(inter[0]), D(inter[1]), None, D(rtol), D(atol), _p(y, D), _p(yk, D), I(nthreads))
    if bad:
        raise RuntimeError(f"oracle_truth_batch_dp: {bad} interval(s) failed")
    return yk


def training_loss(yk, ref, yscale, lb=1e-6, ub=60.0):
    """Per-condition loss_n_ode from knot states yk[N,801,9], labels ref[N,801,7], yscale[N,7] (WIDE_Eoff...:387-396)."""
    pred = np.clip(yk[:, :, :7], lb, ub)
    d = (pred - ref) / yscale[:, None, :]
    return np.mean(d * d, axis=(1, 2))


def set_lb(lb: float = 1.0e-6) -> None:
    """State clamp lower bound used by every C-oracle RHS (process-global; tests restore 1e-6)."""
    lib().oracle_set_lb(ctypes.c_double(lb))


def rhs_batch(T, u, w_in, w_b, w_out, inter=(-30.0, 30.0)):
    T = np.ascontiguousarray(T, np.float64)
    u = np.ascontiguousarray(u, np.float64)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    w_in, w_b, w_out = f(w_in), f(w_b), f(w_out)
    du = np.zeros_like(u)
    F, D, I = ctypes.c_float, ctypes.c_double, ctypes.c_int
    lib().oracle_rhs_batch(I(T.shape[0]), _p(T, D), _p(u, D), _p(w_in, F), _p(w_b, F), _p(w_out, F), D(inter[0]), D(inter[1]), _p(du, D))
    return du
