"""Predictor-MLP training on the device against the reference's own torch operators on the CPU
(oracle/mlp_training_reference.py).  Labels come from a shipped MLP (teacher): the reference's Cantera label files are absent."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


def _teacher_dataset(model_sets, conditions, kind):
    """inputs in physical units and the teacher MLP's un-scaled outputs [n, 800] (torch CPU float32)."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.mlp_training import MlpDataset
    from oracle import reference_path as R
    ms = model_sets("LLNL", "Eon")
    if kind == "temp":
        a = conditions["training_2D"]                         # 800 x (T [K], P [bar])
        mp = R.MLPParams(ms.temp_mlp.w, ms.temp_mlp.b, ms.temp_mlp.out_min, ms.temp_mlp.out_max)
        x = R.scale_inputs([a[:, 0].astype(np.float32), (a[:, 1] * 1e5).astype(np.float32)], 2)
    else:
        a = conditions["independent_4D"]                      # 400 x (T, P [bar], L, u0)
        mp = R.MLPParams(ms.time_mlp.w, ms.time_mlp.b, ms.time_mlp.out_min, ms.time_mlp.out_max)
        x = R.scale_inputs([a[:, 0].astype(np.float32), (a[:, 1] * 1e5).astype(np.float32), a[:, 2].astype(np.float32),
                            a[:, 3].astype(np.float32)], 4)
    y = R.unscale(R.mlp_forward(mp, x), mp.out_min, mp.out_max)
    return MlpDataset(a, y)


def test_dataset_split_and_scaling(model_sets, conditions):
    """800 rows -> 640 / 80 / 80 (the scripts' two train_test_split calls), inputs and outputs in [0, 1]."""
    ds = _teacher_dataset(model_sets, conditions, "temp")
    xt, yt = ds.parts["training"]
    assert xt.shape == (640, 2) and yt.shape == (640, 800) and ds.parts["valid"][0].shape == (80, 2) and ds.parts["test"][1].shape == (80, 800)
    assert xt.dtype == np.float32 and 0.0 <= xt.min() and xt.max() <= 1.0
    allY = np.concatenate([ds.parts[k][1] for k in ("training", "valid", "test")])
    assert abs(allY.min()) < 1e-6 and abs(allY.max() - 1.0) < 1e-6
    from sklearn.model_selection import train_test_split
    idx = np.arange(800)
    tr, te = train_test_split(idx, test_size=0.2, random_state=2024)
    a = conditions["training_2D"]
    assert np.allclose(xt[:, 0], ((a[tr, 0] - 870.0) / 280.0).astype(np.float32))


def test_step_lr_matches_torch_scheduler():
    """The host-side learning-rate schedule equals torch.optim.lr_scheduler.StepLR(100, 0.6) stepped once per epoch."""
    from torch.optim import lr_scheduler
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.mlp_training import TIME_4D_SETTINGS, step_lr
    par = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([par], lr=TIME_4D_SETTINGS.learning_rate)
    sched = lr_scheduler.StepLR(opt, step_size=100, gamma=0.6)
    for epoch in range(450):
        assert abs(step_lr(TIME_4D_SETTINGS, epoch) - opt.param_groups[0]["lr"]) < 1e-15
        opt.step()
        sched.step()


def test_oracle_training_reduces_loss(model_sets, conditions):
    """The oracle itself (reference operators on the CPU): a few epochs on teacher labels bring the loss down."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.mlp_training import TEMP_2D_SETTINGS, epoch_batches, initial_parameters
    from oracle import mlp_training_reference as O
    ds = _teacher_dataset(model_sets, conditions, "temp")
    w, b = initial_parameters(2, seed=1)
    g = torch.Generator().manual_seed(0)
    lists = [(epoch_batches(640, 32, g), epoch_batches(80, 32, g)) for _ in range(3)]
    ht, hv, _ = O.train(ds.parts, TEMP_2D_SETTINGS, w, b, lists, 3)
    assert ht[-1] < 0.25 * ht[0] and all(v >= t for t, v in zip(ht, hv))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["temp", "time"])
def test_device_steps_match_torch_cpu(model_sets, conditions, kind):
    """Thirty optimisation steps on identical mini-batches (incl. a ragged last batch) and a learning-rate change: per-step
    losses agree with torch CPU to 2e-4 relative (the first step to 1e-6; the two float32 trajectories then drift apart
    slowly: GEMM summation orders differ and Adam's division amplifies last-bit differences of small gradients), the
    parameters after the last step to 2e-4 absolute (they have moved by more than 1e-2)."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.mlp_training import MlpTrainSettings, MlpTrainer, initial_parameters
    from oracle import mlp_training_reference as O
    ds = _teacher_dataset(model_sets, conditions, kind)
    x, y = ds.parts["training"]
    in_dim = x.shape[1]
    rng = np.random.default_rng(3)
    sizes = [32] * 28 + [17, 1]
    batches = []
    for bsz in sizes:
        idx = rng.choice(len(x), bsz, replace=False)
        batches.append((x[idx].copy(), y[idx].copy()))
    lrs = [1e-3] * 20 + [6e-4] * 10
    w0, b0 = initial_parameters(in_dim, seed=5)
    ref_loss, ref_w, ref_b = O.run_steps(w0, b0, batches, lrs)
    tr = MlpTrainer(w0, b0, MlpTrainSettings(in_dim=in_dim))
    got = []
    for (bx, by), lr in zip(batches, lrs):
        got.append(tr.step(torch.from_numpy(bx).cuda(), torch.from_numpy(by).cuda(), lr))
    got = torch.stack(got).cpu().numpy()
    assert abs(got[0] - ref_loss[0]) / ref_loss[0] < 1e-6
    assert np.max(np.abs(got - ref_loss) / ref_loss) < 2e-4
    w, b = tr.parameters()
    moved = max(np.max(np.abs(w[i] - w0[i])) for i in range(4))
    err = max(max(np.max(np.abs(w[i] - ref_w[i])), np.max(np.abs(b[i] - ref_b[i]))) for i in range(4))
    print(f"{kind}: loss {ref_loss[0]:.4f} -> {ref_loss[-1]:.5f}, parameters moved {moved:.2e}, max deviation from torch CPU {err:.2e}")
    assert moved > 1e-2 and err < 2e-4
    # eval forward of the trained parameters vs the oracle's model, and determinism of a repeated run
    xe = ds.parts["test"][0]
    out = tr.forward(torch.from_numpy(xe)).cpu().numpy()
    ref_out = O.make_model(ref_w, ref_b)(torch.from_numpy(xe)).detach().numpy()
    assert np.max(np.abs(out - ref_out)) < 2e-4
    tr2 = MlpTrainer(w0, b0, MlpTrainSettings(in_dim=in_dim))
    for (bx, by), lr in zip(batches, lrs):
        tr2.step(torch.from_numpy(bx).cuda(), torch.from_numpy(by).cuda(), lr)
    w2, b2 = tr2.parameters()
    assert all(np.array_equal(w[i], w2[i]) and np.array_equal(b[i], b2[i]) for i in range(4))   # fixed-order reductions


@pytest.mark.gpu
def test_device_training_loop_and_containers(model_sets, conditions, tmp_path):
    """Three epochs of the scripts' loop on the device vs the same loop in torch CPU on the same shuffles: the histories
    (with the scripts' running_loss bookkeeping) agree to 1e-3 relative; the saved .pth / .pkl load through the container
    reader and drive the inference kernels (Surrogate.temp_profile) to the trainer's own predictions."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200 import containers as C
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.mlp_training import (TEMP_2D_SETTINGS, epoch_batches, evaluate_test_set,
                                                                                initial_parameters, train)
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
    from oracle import mlp_training_reference as O
    ds = _teacher_dataset(model_sets, conditions, "temp")
    w0, b0 = initial_parameters(2, seed=2)
    tr, ht, hv = train(ds, TEMP_2D_SETTINGS, w0, b0, num_epochs=3)
    g = torch.Generator().manual_seed(TEMP_2D_SETTINGS.shuffle_seed)
    lists = [(epoch_batches(640, 32, g), epoch_batches(80, 32, g)) for _ in range(3)]
    rt, rv, _ = O.train(ds.parts, TEMP_2D_SETTINGS, w0, b0, lists, 3)
    print("train history", ht, rt, "valid history", hv, rv)
    assert np.allclose(ht, rt, rtol=1e-3) and np.allclose(hv, rv, rtol=1e-3)
    assert ht[-1] < 0.25 * ht[0]
    rep = evaluate_test_set(tr, ds)
    assert rep["accuracy_mean"] > 97.0 and rep["r2"] > 0.5
    pth, pkl = str(tmp_path / "mlp_weights_LLNL_2D_.pth"), str(tmp_path / "min_max_values_mlp_LLNL_2D_.pkl")
    tr.save(pth, pkl, ds.output_scale)
    mp = C.load_mlp(pth, pkl)
    ms = model_sets("LLNL", "Eon")
    sur = Surrogate(C.ModelSet("LLNL", "Eon", ms.crnn, ms.time_mlp, mp), mlp_mode="fp32")
    a = conditions["training_2D"][:64]
    T, P = a[:, 0].astype(np.float32), (a[:, 1] * 1e5).astype(np.float32)
    prof = sur.temp_profile(T, P).cpu().numpy()[1:].T                      # [64, 800] K
    x = np.stack([(a[:, 0] - 870.0) / 280.0, (a[:, 1] - 1.0) / 2.0], 1).astype(np.float32)
    own = tr.forward(torch.from_numpy(x)).double().cpu().numpy() * (ds.output_scale[1] - ds.output_scale[0]) + ds.output_scale[0]
    assert np.max(np.abs(prof - own)) < 5e-3                                # kelvin; float32 un-scaling of ~1000 K values


@pytest.mark.gpu
def test_trainer_argument_validation(model_sets):
    from n_hexane_pyrolysis_surrogate_reactor_model_b200 import _lib
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.mlp_training import MlpTrainSettings, MlpTrainer, initial_parameters
    w0, b0 = initial_parameters(4, seed=0)
    tr = MlpTrainer(w0, b0, MlpTrainSettings(in_dim=4))
    x, y = torch.zeros((33, 4), device="cuda"), torch.zeros((33, 800), device="cuda")
    with pytest.raises(_lib.PfrError):
        tr.step(x, y, 1e-3)                       # more than 32 rows
    with pytest.raises(_lib.PfrError):
        tr.step(x[:8], y[:8], 0.0)                # non-positive learning rate
