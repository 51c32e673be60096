"""Reference-format outputs: prediction dumps, the accuracy table, and training histories the reference can load."""
import numpy as np
import torch

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
