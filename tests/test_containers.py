"""Readers of the reference's container formats, exercised on synthetic files written in those formats, and the
packed fixtures."""
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import packed_path
from n_hexane_pyrolysis_surrogate_reactor_model_b200 import containers as C


def _write_reference_tree(root, mech="LLNL"):
    rng = np.random.default_rng(0)
    os.makedirs(os.path.join(root, "SURROGATE_MODEL_PARAMETER_CONTAINER"))
    os.makedirs(os.path.join(root, "TIME_PRED_MODEL_PARAMETER_CONTAINER"))
    os.makedirs(os.path.join(root, "TEMP_PRED_MODEL_PARAMETER_CONTAINER"))
    hist = {"train_loss": np.array([1.0, 0.5]), "valid_loss": np.array([1.0, 0.6]),
            "parameters": np.array([{"w_in": rng.random((11, 9), np.float32), "w_b": rng.random(9, np.float32),
                                     "w_out": rng.random((9, 9), np.float32)} for _ in range(2)], dtype=object)}
    for key, fname in C.CRNN_FILES.items():
        if key[0] == mech:
            np.savez(os.path.join(root, "SURROGATE_MODEL_PARAMETER_CONTAINER", fname), **hist)

    def mlp(d, in_dim, stem, pk):
        sd = {}
        dims = [(512, in_dim), (512, 512), (512, 512), (800, 512)]
        for i, (o, k) in enumerate(dims, 1):
            sd[f"fc{i}.weight"] = torch.randn(o, k)
            sd[f"fc{i}.bias"] = torch.randn(o)
        torch.save(sd, os.path.join(root, d, stem))
        with open(os.path.join(root, d, pk), "wb") as f:
            pickle.dump({"min": np.float64(0.125), "max": np.float64(3.5)}, f)
        return sd

    for e in ("on", "off"):
        mlp("TIME_PRED_MODEL_PARAMETER_CONTAINER", 4, f"mlp_weights_{mech}_4D_time_{e}.pth", f"min_max_values_mlp_{mech}_4D_time_{e}.pkl")
    mlp("TEMP_PRED_MODEL_PARAMETER_CONTAINER", 2, f"mlp_weights_{mech}_2D.pth", f"min_max_values_mlp_{mech}_2D.pkl")
    return hist


def test_reference_tree_round_trip(tmp_path):
    hist = _write_reference_tree(str(tmp_path))
    on = C.ModelSet.from_reference_dir(str(tmp_path), "LLNL", "Eon")
    off = C.ModelSet.from_reference_dir(str(tmp_path), "LLNL", "Eoff")
    assert on.temp_mlp is not None and off.temp_mlp is None
    assert on.time_mlp.in_dim == 4 and on.temp_mlp.in_dim == 2
    assert (on.time_mlp.out_min, on.time_mlp.out_max) == (0.125, 3.5)
    assert np.array_equal(off.crnn.w_in, hist["parameters"][-1]["w_in"])          # parameters[-1]
    c0 = C.load_npz_parameters(os.path.join(str(tmp_path), "SURROGATE_MODEL_PARAMETER_CONTAINER", C.CRNN_FILES[("LLNL", "Eon")]), epoch=0)
    assert np.array_equal(c0.w_b, hist["parameters"][0]["w_b"])
    out = tmp_path / "packed.npz"
    C.pack_mechanism(str(tmp_path), "LLNL", str(out))
    again = C.ModelSet.from_packed(str(out), "Eon")
    assert all(np.array_equal(a, b) for a, b in zip(again.time_mlp.w + again.temp_mlp.b, on.time_mlp.w + on.temp_mlp.b))
    assert np.array_equal(again.crnn.w_out, on.crnn.w_out)


def test_bad_shapes_are_rejected():
    with pytest.raises(ValueError):
        C.CRNNParams(np.zeros((10, 9)), np.zeros(9), np.zeros((9, 9)))
    with pytest.raises(ValueError):
        C.MLPParams([np.zeros((512, 3)), np.zeros((512, 512)), np.zeros((512, 512)), np.zeros((800, 512))],
                    [np.zeros(512)] * 3 + [np.zeros(800)], 0.0, 1.0)


def test_condition_csv(tmp_path):
    p4 = tmp_path / "c4.csv"
    p4.write_text("1139.4648449138836,1.6938037547233382,0.6405209298008192,4.645800082454905\n"
                  "1117.1135445424156,1.3490843921828986,0.8509428521590036,3.451624433274671\n")
    T, P, L, U = C.load_conditions_csv(str(p4))
    assert T.dtype == np.float32 and T.shape == (2,)
    assert P[0] == np.float32(1.6938037547233382 * 1.0e5)                          # bar -> Pa in float64, then cast
    p2 = tmp_path / "c2.csv"
    p2.write_text("1084.5734992460393,2.836069154391544\n")
    T, P, L, U = C.load_conditions_csv(str(p2))
    assert (float(L[0]), float(U[0])) == (1.0, 2.5)                                # 2-column files: L = 1 m, u0 = 2.5 m/s
    p3 = tmp_path / "c3.csv"
    p3.write_text("1,2,3\n")
    with pytest.raises(ValueError):
        C.load_conditions_csv(str(p3))


@pytest.mark.parametrize("mech", C.MECHANISMS)
def test_packed_fixtures_load(mech):
    for variant in ("Eon", "Eoff"):
        ms = C.ModelSet.from_packed(packed_path(mech), variant)
        assert ms.mechanism == mech and ms.energy_on == (variant == "Eon")
        assert ms.time_mlp.w[3].shape == (800, 512)
        assert 1e-4 < ms.time_mlp.out_min < 2e-4 and 0.3 < ms.time_mlp.out_max < 0.4
        if variant == "Eon":
            assert 860 < ms.temp_mlp.out_min < 870 and 1100 < ms.temp_mlp.out_max < 1140
