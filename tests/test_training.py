"""Training step (config 5): ParameterConverter, loss and adjoint gradient, optimiser step."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import GOLDEN, cond4, packed_path, rel_err


# ----------------------------------------------------------------------------------------------- CPU
def test_product_converter_reproduces_reference_known_answers():
    """The product's ParameterConverter (one projector product instead of nine solves) on the reference's stored
    updated_p -> final_parameters pairs, and against the oracle's literal restatement."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import NARROW, ConverterSpec, ParameterConverter, narrow_spec
    from oracle import reference_path as R
    kat = np.load(f"{GOLDEN}/converter_kat.npz")
    for name, spec, ospec in (("LLNL_Eoff_wide", ConverterSpec(), R.ConverterSpec()),
                              ("NUIG_Eon", ConverterSpec(form="eon", b_fit=1.858, Ea_fit=58.397, **NARROW), R.narrow_spec("eon", 1.858, 58.397)),
                              ("NUIG_Eon", narrow_spec("eon", "NUIG"), R.narrow_spec("eon", 1.858, 58.397))):
        p = torch.tensor(kat[f"{name}/updated_p"])
        w_in, w_b, w_out = ParameterConverter(spec)(p)
        assert np.max(np.abs(w_in.numpy() - kat[f"{name}/w_in"])) < 3e-6
        assert np.array_equal(w_b.numpy(), kat[f"{name}/w_b"])
        assert np.max(np.abs(w_out.numpy() - kat[f"{name}/w_out"])) < 3e-6
        o_in, o_b, o_out = R.parameter_converter(p, ospec)
        assert torch.allclose(w_in, o_in, atol=3e-6) and torch.allclose(w_out, o_out, atol=3e-6) and torch.equal(w_b, o_b)


def test_converter_gradient_has_the_reference_dead_slice():
    """p[108:189] never reaches the outputs (overwritten by clamp(-w_out), WIDE_Eoff...:200,221): zero gradient."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import ParameterConverter
    p = (torch.rand(189) * 0.8 + 0.2).requires_grad_(True)
    w_in, w_b, w_out = ParameterConverter()(p)
    (w_in.sum() + w_b.sum() + (w_out ** 2).sum()).backward()
    assert torch.all(p.grad[108:] == 0) and torch.any(p.grad[:108] != 0)


def test_split_indices_and_batch_subsets():
    """The 80/10/10 split the trainers draw with sklearn (random_state 42) and TrainingBatch.subset column picks."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import TrainingBatch, narrow_spec, split_indices
    tr, va, te = split_indices(800)
    assert (len(tr), len(va), len(te)) == (640, 80, 80)
    assert sorted(np.concatenate([tr, va, te]).tolist()) == list(range(800))
    tr2, _, _ = split_indices(800)
    assert np.array_equal(tr, tr2)
    n = 6
    b = TrainingBatch(torch.arange(n, dtype=torch.float32), torch.arange(n, dtype=torch.float32) + 10,
                      torch.arange(801 * n, dtype=torch.float32).view(801, n), torch.arange(801 * n, dtype=torch.float32).view(801, n) + 1,
                      torch.arange(801 * 7 * n, dtype=torch.float32).view(801, 7, n), torch.arange(7 * n, dtype=torch.float32).view(7, n))
    s = b.subset([4, 1])
    assert s.n == 2 and s.T0.tolist() == [4.0, 1.0] and s.c0.tolist() == [14.0, 11.0]
    assert torch.equal(s.tgrid, b.tgrid[:, [4, 1]]) and torch.equal(s.Tprof, b.Tprof[:, [4, 1]])
    assert torch.equal(s.ref, b.ref[:, :, [4, 1]]) and torch.equal(s.yscale, b.yscale[:, [4, 1]]) and s.ref.is_contiguous()
    spec = narrow_spec("eoff", "JetSurf")
    assert (spec.form, spec.b_fit, spec.Ea_fit, spec.wout, spec.A) == ("eoff", 2.1133, 61.713, (-2.0, 2.0), (3.0, 21.0))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ar_worker(rank, world, port, ret):
    import torch.distributed as dist
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import allreduce_packed
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    packed = torch.cat([torch.full((189,), float(rank + 1), dtype=torch.float64), torch.tensor([0.5 * (rank + 1), 320.0], dtype=torch.float64)])
    out = allreduce_packed(packed)
    ret[rank] = bool(torch.all(out[:189] == 3.0)) and float(out[189]) == 1.5 and float(out[190]) == 640.0
    dist.destroy_process_group()


def test_gradient_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_ar_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert all(ret[r] for r in range(2))


# ----------------------------------------------------------------------------------------------- GPU
def _setup(surrogates, model_sets, conditions, n=12):
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import CrnnTrainer, synthetic_labels
    a = conditions["training_wide_2D"]
    sel = np.linspace(0, len(a) - 1, n).astype(int)
    T, P = a[sel, 0].astype(np.float32), (a[sel, 1] * 1e5).astype(np.float32)
    sur = surrogates("LLNL", "Eoff")
    teacher = model_sets("LLNL", "Eoff", "Eoff_wide").crnn
    batch = synthetic_labels(sur, teacher, T, P)
    return CrnnTrainer(batch), batch, T, P


@pytest.mark.gpu
@pytest.mark.parametrize("substeps,forward", [(1, "dp54w"), (2, "dp54w"), (1, "bs23w")])
def test_loss_and_gradient_vs_oracle_finite_differences(surrogates, model_sets, conditions, substeps, forward):
    """Student = the stored LLNL_Eoff_wide_v2 parameters, labels from the LLNL_Eoff_wide teacher, training RHS clamps
    (exponent +-10).  The GPU loss equals the oracle's (converged knot states -> numpy loss) to 1e-7 relative with the
    forward pass at 1e-10, and the adjoint gradient matches central finite differences of the oracle loss taken in
    float64 parameters to 2e-4 of the gradient's scale (measured: 5.6e-6 with one RK4 sub-step of the adjoint per knot interval,
    the trainer's default; 1.2e-6 with two), with either forward pass: free Dormand-Prince steps + continuous extension
    (dp54w, the default of the isothermal trainers) or one BS23 step per knot interval (bs23w)."""
    from oracle import c_oracle as CO
    tr, batch, T, P = _setup(surrogates, model_sets, conditions)
    tr.rtol, tr.atol = 1e-10, 1e-12
    tr.substeps, tr.forward_method = substeps, forward
    student = model_sets("LLNL", "Eoff").crnn
    lsum, gsum, bad = tr.loss_grad_w(student.w_in, student.w_b, student.w_out)
    assert bad == 0
    n = batch.n
    tg = batch.tgrid.cpu().numpy().T.copy()
    Tp = np.repeat(T[:, None], 801, 1)
    c0 = np.zeros((n, 9), np.float32)
    c0[:, 6] = batch.c0.cpu().numpy()
    ref = batch.ref.cpu().numpy().transpose(2, 0, 1).astype(np.float64)      # [n, 801, 7]
    ysc = batch.yscale.cpu().numpy().T.astype(np.float64)                     # [n, 7]

    def oracle_loss(w_in, w_b, w_out):
        yk = CO.truth_knots_dp(tg, Tp, c0, w_in, w_b, w_out, inter=(-10.0, 10.0), nthreads=8)
        return CO.training_loss(yk, ref, ysc).sum()

    w = [student.w_in.astype(np.float64), student.w_b.astype(np.float64), student.w_out.astype(np.float64)]
    L0 = oracle_loss(*w)
    assert abs(float(lsum) - L0) / L0 < 1e-7
    g = gsum.cpu().numpy()
    g_in, g_b, g_out = g[:99].reshape(11, 9), g[99:108], g[108:].reshape(9, 9)
    scale = np.abs(g).max()
    checks = [(0, (6, 0)), (0, (0, 2)), (0, (9, 0)), (0, (9, 4)), (0, (10, 0)), (0, (10, 6)), (1, (0,)), (1, (3,)), (2, (6, 0)),
              (2, (2, 0)), (2, (0, 4)), (2, (8, 8))]
    worst = 0.0
    for which, idx in checks:
        h = 1e-6 * max(1.0, abs(w[which][idx]))
        wp = [a.copy() for a in w]
        wm = [a.copy() for a in w]
        wp[which][idx] += h
        wm[which][idx] -= h
        fd = (oracle_loss(*wp) - oracle_loss(*wm)) / (2 * h)
        got = (g_in, g_b, g_out)[which][idx]
        worst = max(worst, abs(got - fd) / scale)
        assert abs(got - fd) < 2e-4 * scale + 1e-6 * abs(fd), (which, idx, got, fd)
    print(f"adjoint gradient vs finite differences, {tr.substeps} sub-step(s), forward {forward}: worst deviation {worst:.2e} of the gradient scale")


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["Eoff", "Eon"])
def test_three_adjoint_implementations_agree(surrogates, model_sets, conditions, variant):
    """The three-kernel adjoint (product: node matrices in parallel, sequential walk of 9 x 9 mat-vecs, parallel gradient
    quadrature), the one-condition-per-warp kernel and the one-condition-per-thread kernel are the same discretisation with
    different mappings and summation orders: losses agree to 1e-12, all 189 gradient entries to 1e-10 of the gradient scale --
    isothermal (wide trainer) and along a temperature ramp (narrow Eon trainer), ragged batch size."""
    if variant == "Eoff":
        tr, batch, T, P = _setup(surrogates, model_sets, conditions, n=9)
        student = model_sets("LLNL", "Eoff").crnn
    else:
        tr, batch, T, P = _setup_eon(surrogates, model_sets, conditions, n=7)
        student = model_sets("JetSurf", "Eon").crnn
    out = {}
    for name, adjoint, sub in (("staged", "staged", 2), ("warp", "warp", 2), ("thread", "warp", -2), ("staged3", "staged", 3), ("warp3", "warp", 3)):
        tr.adjoint, tr.substeps = adjoint, sub
        l, g, bad = tr.loss_grad_w(student.w_in, student.w_b, student.w_out)
        assert bad == 0
        out[name] = (float(l), g.clone())
    for a, b in (("staged", "warp"), ("warp", "thread"), ("staged3", "warp3")):
        assert abs(out[a][0] - out[b][0]) < 1e-12 * abs(out[b][0]), (a, b)
        assert float((out[a][1] - out[b][1]).abs().max()) < 1e-10 * float(out[b][1].abs().max()), (a, b)


@pytest.mark.gpu
def test_gradient_vs_reference_style_autograd(surrogates, model_sets, conditions):
    """The reference obtains its gradient by back-propagating through torchdiffeq's dopri5 operations (float32,
    rtol 1e-4, atol 1e-6: discretise-then-differentiate).  The adjoint gradient of the SAME loss at the same
    tolerances agrees with autograd through the oracle's dopri5 restatement to within that solver tolerance level:
    5 % of the gradient scale per block (measured ~1 %), loss value to 1e-3 relative."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import CrnnTrainer, TrainingBatch
    from oracle import reference_path as R
    torch.set_num_threads(1)
    tr, batch, T, P = _setup(surrogates, model_sets, conditions, n=3)
    student = model_sets("LLNL", "Eoff").crnn
    for i in range(batch.n):
        one = TrainingBatch(batch.T0[i:i + 1].contiguous(), batch.c0[i:i + 1].contiguous(), batch.tgrid[:, i:i + 1].contiguous(), None,
                            batch.ref[:, :, i:i + 1].contiguous(), batch.yscale[:, i:i + 1].contiguous())
        lsum, gsum, bad = CrnnTrainer(one).loss_grad_w(student.w_in, student.w_b, student.w_out)
        assert bad == 0
        g = gsum.cpu().numpy()
        w_in = torch.tensor(student.w_in, requires_grad=True)
        w_b = torch.tensor(student.w_b, requires_grad=True)
        w_out = torch.tensor(student.w_out, requires_grad=True)
        tt = one.tgrid[:, 0].cpu()
        f = R.CRNNFunc(tt, torch.full((801,), float(T[i])), w_in, w_b, w_out, inter=(-10.0, 10.0))
        u0 = torch.zeros(9)
        u0[6] = one.c0[0].cpu()
        sol = R.odeint_dopri5(f, u0, tt, rtol=1e-4, atol=1e-6)
        ref = torch.cat([one.ref[:, :, 0].cpu().T, torch.zeros(2, 801)])
        ysc = torch.cat([one.yscale[:, 0].cpu(), torch.ones(2)])
        loss = R.loss_n_ode(torch.clamp(sol.T, 1e-6, 60.0), ref, ysc)
        loss.backward()
        assert abs(float(lsum) - float(loss)) / float(loss) < 1e-3
        for got, want in ((g[:99].reshape(11, 9), w_in.grad.numpy()), (g[99:108], w_b.grad.numpy()), (g[108:].reshape(9, 9), w_out.grad.numpy())):
            assert np.max(np.abs(got - want)) < 0.05 * np.abs(want).max()


@pytest.mark.gpu
def test_training_steps_reduce_the_loss(surrogates, model_sets, conditions):
    """A few reference-style steps (clip 10, AdamW 5e-4, wd 1e-4) from the stored updated_p perturbed with N(0, 0.05^2)
    noise (SURVEY config 5): the loss goes down and no trajectory fails."""
    kat = np.load(f"{GOLDEN}/converter_kat.npz")
    tr, batch, T, P = _setup(surrogates, model_sets, conditions, n=32)
    g = torch.Generator().manual_seed(0)
    p = (torch.tensor(kat["LLNL_Eoff_wide/updated_p"]) + 0.05 * torch.randn(189, generator=g)).requires_grad_(True)
    losses = []
    for _ in range(6):
        loss, bad = tr.step(p)
        assert bad == 0 and np.isfinite(loss)
        losses.append(loss)
    assert losses[-1] < losses[0]


# ----------------------------------------------------------------------------------------------- narrow Eon trainer
def _setup_eon(surrogates, model_sets, conditions, n=8, mech="NUIG"):
    """Eon_surrogate_model_training.py's set-up with synthetic labels: temperature profile from the temperature MLP,
    teacher = the stored <mech>_Eon CRNN, narrow RHS clamps (lb 1e-5, exponent +-30), converter form 'eon'."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import TRAINING_NARROW_CLAMPS
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import NARROW_EON_SETTINGS, CrnnTrainer, narrow_spec, synthetic_labels
    a = conditions["training_2D"]
    sel = np.linspace(0, len(a) - 1, n).astype(int)
    T, P = a[sel, 0].astype(np.float32), (a[sel, 1] * 1e5).astype(np.float32)
    sur = surrogates(mech, "Eon")
    batch = synthetic_labels(sur, model_sets(mech, "Eon").crnn, T, P, clamps=TRAINING_NARROW_CLAMPS)
    assert batch.Tprof is not None
    return CrnnTrainer(batch, spec=narrow_spec("eon", mech), settings=NARROW_EON_SETTINGS), batch, T, P


@pytest.mark.gpu
def test_eon_loss_and_gradient_vs_oracle_finite_differences(surrogates, model_sets, conditions):
    """The narrow Eon trainer's loss and gradient along MLP-predicted temperature ramps (adjoint with dT/dt terms):
    student = JetSurf_Eon parameters against NUIG_Eon labels.  Loss to 1e-7 of the oracle's converged value, gradient to
    2e-4 of its scale against float64 central differences of the oracle loss (state clamp lower bound 1e-5)."""
    from oracle import c_oracle as CO
    tr, batch, T, P = _setup_eon(surrogates, model_sets, conditions, n=6)
    tr.rtol, tr.atol = 1e-10, 1e-12
    student = model_sets("JetSurf", "Eon").crnn
    lsum, gsum, bad = tr.loss_grad_w(student.w_in, student.w_b, student.w_out)
    assert bad == 0
    n = batch.n
    tg = batch.tgrid.cpu().numpy().T.copy()
    Tp = batch.Tprof.cpu().numpy().T.copy()
    c0 = np.zeros((n, 9), np.float32)
    c0[:, 6] = batch.c0.cpu().numpy()
    ref = batch.ref.cpu().numpy().transpose(2, 0, 1).astype(np.float64)
    ysc = batch.yscale.cpu().numpy().T.astype(np.float64)

    def oracle_loss(w_in, w_b, w_out):
        yk = CO.truth_knots_dp(tg, Tp, c0, w_in, w_b, w_out, inter=(-30.0, 30.0), nthreads=8)
        return CO.training_loss(yk, ref, ysc, lb=1e-5).sum()

    CO.set_lb(1e-5)
    try:
        w = [student.w_in.astype(np.float64), student.w_b.astype(np.float64), student.w_out.astype(np.float64)]
        L0 = oracle_loss(*w)
        assert abs(float(lsum) - L0) / L0 < 1e-7
        g = gsum.cpu().numpy()
        g_in, g_b, g_out = g[:99].reshape(11, 9), g[99:108], g[108:].reshape(9, 9)
        scale = np.abs(g).max()
        for which, idx in [(0, (6, 0)), (0, (9, 0)), (0, (9, 5)), (0, (10, 0)), (0, (10, 3)), (1, (0,)), (1, (7,)), (2, (6, 0)), (2, (1, 2)), (2, (4, 8))]:
            h = 1e-6 * max(1.0, abs(w[which][idx]))
            wp = [a.copy() for a in w]
            wm = [a.copy() for a in w]
            wp[which][idx] += h
            wm[which][idx] -= h
            fd = (oracle_loss(*wp) - oracle_loss(*wm)) / (2 * h)
            got = (g_in, g_b, g_out)[which][idx]
            assert abs(got - fd) < 2e-4 * scale + 1e-6 * abs(fd), (which, idx, got, fd)
    finally:
        CO.set_lb(1e-6)


@pytest.mark.gpu
def test_minibatch_schedule_and_eon_training_loop(surrogates, model_sets, conditions, tmp_path):
    """train() on the narrow Eon trainer: (i) one un-shuffled epoch with batch_size = n is exactly one full-batch step;
    (ii) the reference's per-sample schedule (batch_size 1, shuffled) and mini-batches of 4 both lower the validation
    loss from the perturbed stored parameters; (iii) the history file is written in the reference's layout."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import NARROW_EON_SETTINGS, CrnnTrainer, narrow_spec, train
    kat = np.load(f"{GOLDEN}/converter_kat.npz")
    tr, batch, T, P = _setup_eon(surrogates, model_sets, conditions, n=8)
    g = torch.Generator().manual_seed(1)
    p0 = torch.tensor(kat["NUIG_Eon/updated_p"]) + 0.03 * torch.randn(189, generator=g)
    mk = lambda: CrnnTrainer(batch, spec=narrow_spec("eon", "NUIG"), settings=NARROW_EON_SETTINGS)
    pa, pb = p0.clone().requires_grad_(True), p0.clone().requires_grad_(True)
    train(mk(), pa, 1, batch_size=batch.n, shuffle_seed=None)
    mk().step(pb)
    assert torch.equal(pa.detach(), pb.detach())
    for bs in (1, 4):
        p = p0.clone().requires_grad_(True)
        path = str(tmp_path / f"hist_{bs}.npz")
        h = train(mk(), p, 3, batch_size=bs, save_path=path)
        first = mk().loss(p0)
        assert abs(first - mk().loss_and_grad(p0)[0]) < 1e-12 * first
        assert h["valid_loss"][-1] < first, (bs, first, h["valid_loss"])
        z = np.load(path, allow_pickle=True)
        assert len(z["train_loss"]) == 3 and z["parameters"][-1]["w_out"].shape == (9, 9) and z["updated_p"].shape == (189,)


# ----------------------------------------------------------------------------------------------- two GPUs, NCCL
def _nccl_worker(rank, world, port, ret):
    """Each rank: its contiguous half of the 12-condition batch -> packed [grad | loss | ok | attempted] -> NCCL all-reduce."""
    import torch.distributed as dist
    from conftest import GOLDEN, packed_path
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import shard_bounds
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import CrnnTrainer, allreduce_packed, synthetic_labels
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    a = np.load(os.path.join(GOLDEN, "conditions.npz"))["training_wide_2D"]
    sel = np.linspace(0, len(a) - 1, 12).astype(int)
    lo, hi = shard_bounds(len(sel), world, rank)
    T, P = a[sel[lo:hi], 0].astype(np.float32), (a[sel[lo:hi], 1] * 1e5).astype(np.float32)
    sur = Surrogate(ModelSet.from_packed(packed_path("LLNL"), "Eoff"), device=f"cuda:{rank}")
    teacher = ModelSet.from_packed(packed_path("LLNL"), "Eoff", "Eoff_wide").crnn
    student = ModelSet.from_packed(packed_path("LLNL"), "Eoff").crnn
    tr = CrnnTrainer(synthetic_labels(sur, teacher, T, P))
    packed = allreduce_packed(tr.packed_loss_grad(student.w_in, student.w_b, student.w_out))
    ret[rank] = packed.cpu().numpy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_gradient_allreduce_equals_single_gpu(surrogates, model_sets, conditions):
    """The training step's only collective on real hardware: two ranks, each with half of the batch, all-reduce their packed
    vectors over NCCL; the result must equal the single-GPU reduction over the whole batch (both sum the same per-condition rows in
    float64; the summation trees differ, hence 1e-12 relative to the vector's scale instead of bit equality)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    tr, batch, T, P = _setup(surrogates, model_sets, conditions)
    student = model_sets("LLNL", "Eoff").crnn
    single = tr.packed_loss_grad(student.w_in, student.w_b, student.w_out).cpu().numpy()
    ret = mp.Manager().dict()
    mp.spawn(_nccl_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    scale = np.abs(single[:189]).max()
    for r in range(2):
        assert np.max(np.abs(ret[r][:189] - single[:189])) <= 1e-12 * scale, r
        assert abs(ret[r][189] - single[189]) <= 1e-12 * abs(single[189])
        assert ret[r][190] == single[190] == 12.0 and ret[r][191] == 12.0
    assert np.array_equal(ret[0], ret[1])
