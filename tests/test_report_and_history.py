"""Reference-format outputs: prediction dumps, the accuracy table, and training histories the reference can load."""
import numpy as np
import pytest
import torch

from conftest import cond4

from n_hexane_pyrolysis_surrogate_reactor_model_b200 import report
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import load_npz_parameters


def test_accuracy_rows_match_the_reference_formulas():
    from oracle import reference_path as R
    rng = np.random.default_rng(1)
    true = np.abs(rng.normal(1.0, 0.5, (7, 60))).astype(np.float32)
    pred = (true * (1 + 0.05 * rng.normal(size=true.shape))).astype(np.float32)
    for absden in (False, True):
        rows = report.accuracy_rows(3, pred, true, 1000.0, 2e5, 0.7, 3.0, absolute_denominator=absden)
        assert [r[1] for r in rows] == report.SPECIES_OBS and rows[0][:6] == [3, "H2", 1000.0, 2e5, 0.7, 3.0]
        for s in range(7):
            want = R.accuracy_metrics(pred[s], true[s], absden)
            assert np.allclose(rows[s][6:], want, rtol=1e-6, atol=0)
    df = report.accuracy_table(rows)
    assert list(df.columns) == report.COLUMNS and len(df) == 7


def test_prediction_files_have_the_reference_layout(tmp_path):
    n, nt = 3, 801
    rng = np.random.default_rng(2)
    tgrid = np.cumsum(rng.uniform(1e-5, 1e-3, (nt, n)), axis=0).astype(np.float32)
    tgrid[0] = 0
    Tprof = rng.uniform(900, 1100, (nt, n)).astype(np.float32)
    dense = rng.uniform(0.1, 5.0, (nt, 9, n))
    T, P, L, U = Tprof[0], np.full(n, 2e5, np.float32), np.full(n, 0.7, np.float32), np.full(n, 3.0, np.float32)
    idx = np.array([800, 278, 10])
    paths = report.write_prediction_files(str(tmp_path), "pred_LLNLon_", tgrid, Tprof, dense, T, P, L, U, idx_cut=idx)
    assert [p.split("/")[-1] for p in paths] == ["pred_LLNLon_1.txt", "pred_LLNLon_2.txt", "pred_LLNLon_3.txt"]
    a = np.loadtxt(paths[1])
    assert a.shape == (279, 12)                                   # [:idx_cut + 1] rows, [t, T, P, L, u0, 7 species]
    assert np.all(a[0, 5:11] == 0.0) and a[0, 11] > 0             # products forced to 0 at t = 0, n-hexane kept
    assert np.allclose(a[:, 0], tgrid[:279, 1], rtol=1e-6) and np.allclose(a[5, 5:], dense[5, :7, 1], rtol=1e-6)
    assert open(paths[0]).readline().split()[0] == "0.000000e+00"  # fmt="%.6e"
    iso = report.write_prediction_files(str(tmp_path / "off"), "pred_LLNLoff_", tgrid, None, dense, T, P, L, U)
    assert np.loadtxt(iso[0]).shape == (801, 12) and np.allclose(np.loadtxt(iso[0])[:, 1], T[0], rtol=1e-6)   # '%.6e' keeps 7 digits


def test_nearest_time_labels():
    t_label = np.array([0.0, 1.0, 2.0, 3.0])
    y = np.arange(8.0).reshape(2, 4)
    got = report.nearest_time_labels([0.4, 0.6, 2.9], t_label, y)
    assert np.array_equal(got, y[:, [0, 1, 3]])


def test_history_files_load_like_the_reference_ones(tmp_path):
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import ParameterConverter, save_history
    conv = ParameterConverter()
    hist = {"train_loss": [], "valid_loss": [], "parameters": []}
    p = torch.rand(189) * 0.8 + 0.2
    for e in range(3):
        w = [x.numpy() for x in conv(p + 0.01 * e)]
        hist["train_loss"].append(1.0 / (e + 1))
        hist["valid_loss"].append(1.1 / (e + 1))
        hist["parameters"].append({"w_in": w[0], "w_b": w[1], "w_out": w[2]})
    path = str(tmp_path / "training_history_test.npz")
    save_history(path, hist, final=tuple(w), p=p)
    d = np.load(path, allow_pickle=True)
    assert set(d.files) == {"train_loss", "valid_loss", "parameters", "final_parameters", "updated_p"}
    assert d["parameters"].shape == (3,) and d["updated_p"].shape == (189,) and d["updated_p"].dtype == np.float32
    last = load_npz_parameters(path)                               # the reference's loader semantics: parameters[-1]
    assert np.array_equal(last.w_in, w[0]) and np.array_equal(d["final_parameters"].item()["w_out"], w[2])


def test_three_condition_pick_follows_the_reference():
    """sorted(test_idx, key=T)[n//4], [n//2], [-2]   (...Eoff_validation_plot.py:367-373)."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.validation import three_conditions
    T = np.array([900.0, 1100.0, 950.0, 1000.0, 875.0, 1150.0, 1050.0, 980.0])
    test_idx = [7, 1, 4, 5, 2, 0]                      # sorted by T: 4(875) 0(900) 2(950) 7(980) 1(1100) 5(1150)
    assert three_conditions(test_idx, T) == [0, 7, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["Eoff", "Eon"])
def test_two_model_comparison_driver(surrogates, conditions, variant, tmp_path):
    """Both mechanisms through the same kernels at the three picked conditions; tables in the reference's prediction
    layout, equal to what predict_n_ode gives for those conditions, Eon trimmed at idx_cut."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.validation import two_model_comparison
    T, P, L, U = cond4(conditions, n=40)
    s1, s2 = surrogates("LLNL", variant), surrogates("NUIG", variant)
    out = two_model_comparison(s1, s2, T, P, L, U, out_dir=str(tmp_path), names=("LLNL", "NUIG"))
    idx = out["conditions"]
    assert len(idx) == 3 and T[idx[0]] <= T[idx[1]] <= T[idx[2]]
    for name, sur in (("LLNL", s1), ("NUIG", s2)):
        for j, i in enumerate(idx):
            tab = out[name][j]
            assert tab.shape[1] == 12 and tab[0, 5:11].max() == 0.0 and tab[0, 11] > 0
            if variant == "Eon":
                res = sur.predict_n_ode(T[i:i + 1], P[i:i + 1])
                _, tend = sur.time_grid(T[i:i + 1], P[i:i + 1], L[i:i + 1], U[i:i + 1], want_grid=False, want_end=True)
                k = int(sur.idx_cut(res.tgrid, tend)[0]) + 1
            else:
                res = sur.predict_n_ode(T[i:i + 1], P[i:i + 1], L[i:i + 1], U[i:i + 1])
                k = 801
            assert len(tab) == k
            assert np.allclose(tab[1:, 5:], res.dense[1:k, :7, 0].cpu().numpy(), rtol=2e-6, atol=1e-7)
            saved = np.loadtxt(tmp_path / f"{name}_cond{i + 1}.txt")
            assert np.allclose(saved, tab, rtol=1e-6, atol=1e-12)
    assert len(open(tmp_path / "comparison_index.csv").read().splitlines()) == 7


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["Eoff", "Eon"])
def test_device_accuracy_kernel_matches_reference_formulas(surrogates, model_sets, conditions, variant):
    """pfr_accuracy over a whole batch against the oracle's per-species restatement of the reference formulas
    (Eoff: ref + eps denominators, all 800 knots; Eon: |ref| + eps, trimmed at idx_cut).  Labels = another mechanism's
    trajectories so that every metric is non-trivial.  1e-5 relative (the reference reduces in float32)."""
    import torch
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.report import accuracy_device, accuracy_rows, accuracy_rows_device
    from oracle import reference_path as R
    T, P, L, U = cond4(conditions, n=24)
    sur, other = surrogates("LLNL", variant), surrogates("JetSurf", variant)
    res = sur.predict_n_ode(T, P) if variant == "Eon" else sur.predict_n_ode(T, P, L, U)
    helper = type(sur).__new__(type(sur))
    helper.device, helper.energy_on, helper.crnn = sur.device, sur.energy_on, other.crnn
    lab = helper.integrate(torch.as_tensor(T), sur.inlet_concentration(T, P), tgrid=res.tgrid, Tprof=res.Tprof, dense=True)
    labels = lab.dense[:, :7, :].to(torch.float32).contiguous()
    idx = None
    if variant == "Eon":
        _, tend = sur.time_grid(T, P, L, U, want_grid=False, want_end=True)
        idx = sur.idx_cut(res.tgrid, tend)
    m = accuracy_device(res.dense, labels, idx, absolute_denominator=(variant == "Eon")).cpu().numpy()
    pred = res.dense.cpu().numpy().astype(np.float32)
    true = labels.cpu().numpy()
    for i in range(len(T)):
        k = 801 if idx is None else int(idx[i]) + 1
        for s in range(7):
            want = R.accuracy_metrics(pred[:k, s, i].astype(np.float64), true[:k, s, i].astype(np.float64), variant == "Eon")
            got = m[:, s, i]
            assert np.allclose(got, want, rtol=1e-5, atol=1e-9), (i, s, got, want)
    rows = accuracy_rows_device(res.dense, labels, T, P, L, U, idx, variant == "Eon")
    host = accuracy_rows(1, pred[:(801 if idx is None else int(idx[0]) + 1), :7, 0].T, true[:(801 if idx is None else int(idx[0]) + 1), :, 0].T,
                         T[0], P[0], L[0], U[0], absolute_denominator=(variant == "Eon"))
    assert len(rows) == 7 * len(T) and rows[0][:2] == [1, "H2"]
    assert np.allclose(np.array([r[6:] for r in rows[:7]], float), np.array([r[6:] for r in host], float), rtol=2e-4, atol=1e-7)
