"""One CRNN training step on the GPU, batched over conditions and sharded over ranks.

Mirrors SURROGATE_MODEL_TRAINING/WIDE_Eoff_surrogate_model_training.py (and its narrow Eoff / Eon siblings):

  ParameterConverter(p)            :194-228   flat p[189] -> (w_in[11,9], w_b[9], w_out[9,9]); element-balance
                                              projection of w_out, clamps, w_in[:9] = clamp(-w_out, 0, ul)
  Trainer.predict_n_ode(p, i_exp)  :370-385   ParameterConverter -> odeint(rtol 1e-4, atol 1e-6) -> clamp
  Trainer.loss_n_ode(p, i_exp)     :387-396   MSE of prediction / yscale vs label / yscale over 7 species x 801 points
  loss.backward(); clip_grad_norm_([p], 10); AdamW(lr 5e-4, wd 1e-4).step()   :414-422, :500

The reference updates after every single sample (batch size 1, ~0.23 s per sample on a CPU core).  Here one step
takes the whole local shard of conditions at once: forward trajectories and the adjoint gradient are two kernel
launches of the C-ABI library (pfr_integrate with dense raw output, pfr_loss_grad), the tiny ParameterConverter
and its backward run as torch autograd on 189 numbers, and the only communication is one all-reduce of
[grad(189) | loss | count] (NCCL on GPUs, gloo in the CPU test).
"""
from __future__ import annotations

import os

import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .containers import CRNNParams
from .surrogate import NS, NTOTAL, TRAINING_NARROW_CLAMPS, TRAINING_WIDE_CLAMPS, CrnnModel, Surrogate, _ptr, _stream

NR = 9
NPAR = 189
E_H = (2, 4, 4, 6, 6, 8, 14, 10, 10)   # WIDE_Eoff_surrogate_model_training.py:129-130
E_C = (0, 1, 2, 2, 3, 4, 6, 4, 5)


@dataclass
class ConverterSpec:
    """Fits, clamps and slope formulas of one trainer (form 'wide' | 'eon' | 'eoff')."""
    form: str = "wide"
    A_fit: float = 18.42068
    b_fit: float = 2.112
    Ea_fit: float = 63.304
    slope_reg: float = 0.5
    wout: tuple = (-5.0, 5.0)
    win: tuple = (0.0, 5.0)
    Ea: tuple = (5.0, 200.0)
    b: tuple = (-3.0, 3.0)
    A: tuple = (1.0, 21.0)


WIDE_LLNL = ConverterSpec()                                                    # WIDE_Eoff...:25-29,48-52
NARROW = dict(wout=(-2.0, 2.0), win=(0.0, 2.0), Ea=(10.0, 200.0), b=(-3.0, 3.0), A=(3.0, 21.0))   # Eon...:51-56
NARROW_FITS = {"LLNL": (2.3263, 67.933), "NUIG": (1.858, 58.397), "JetSurf": (2.1133, 61.713)}     # (b_fit, Ea_fit)  Eon...:31-40


def narrow_spec(form: str, mechanism: str = "LLNL") -> ConverterSpec:
    """ConverterSpec of the narrow-range trainers: form 'eon' (Eon_surrogate_model_training.py:287-327) or 'eoff'
    (Eoff_surrogate_model_training.py:204-244) with the mechanism's Arrhenius fits."""
    b_fit, Ea_fit = NARROW_FITS[mechanism]
    return ConverterSpec(form=form, b_fit=b_fit, Ea_fit=Ea_fit, **NARROW)


@dataclass
class TrainerSettings:
    """Solver tolerances, RHS clamps and optimiser constants of one reference trainer script."""
    clamps: tuple
    rtol: float
    atol: float
    lr: float
    weight_decay: float
    clip: float
    lr_factor: float      # ReduceLROnPlateau factor (patience 5, threshold 1e-4 rel everywhere)


WIDE_EOFF_SETTINGS = TrainerSettings(TRAINING_WIDE_CLAMPS, 1e-4, 1e-6, 5e-4, 1e-4, 10.0, 0.8)        # WIDE_Eoff...:16-19,383,500-501
NARROW_EOFF_SETTINGS = TrainerSettings(TRAINING_NARROW_CLAMPS, 1e-2, 1e-3, 5e-3, 1e-2, 10.0, 0.6)    # Eoff...:20,397,514-515 (AdamW default wd)
NARROW_EON_SETTINGS = TrainerSettings(TRAINING_NARROW_CLAMPS, 1e-2, 1e-3, 5e-3, 1e-2, 10.0, 0.5)     # Eon...:22,480,597-598


def split_indices(n_exp: int, seed: int = 42):
    """train / valid / test condition indices exactly as the trainers draw them (WIDE_Eoff...:57-58, Eon...:61-62):
    sklearn train_test_split 80/20 then 50/50 of the remainder, random_state 42."""
    from sklearn.model_selection import train_test_split
    train_idx, temp_idx = train_test_split(np.arange(n_exp), test_size=0.2, random_state=seed)
    valid_idx, test_idx = train_test_split(temp_idx, test_size=0.5, random_state=seed)
    return train_idx, valid_idx, test_idx


class ParameterConverter:
    """p[189] -> (w_in, w_b, w_out), differentiable (torch autograd), float32 like the reference."""

    def __init__(self, spec: ConverterSpec = WIDE_LLNL, device="cpu"):
        self.spec = spec
        f32 = torch.float32
        E_ = torch.stack([torch.tensor(E_H, dtype=f32), torch.tensor(E_C, dtype=f32)], dim=1)
        _, _, Vh = torch.linalg.svd(E_.T, full_matrices=True)                  # :132-133
        self.E_null = Vh[E_.size(1):].T.contiguous().to(device)
        A, b, Ea, reg = (torch.tensor(v, dtype=f32) for v in (spec.A_fit, spec.b_fit, spec.Ea_fit, spec.slope_reg))
        if spec.form == "wide":                                                # :186-188
            sA, sb, sE = A * (A / (A + NR)) * reg, b * ((A + b + NR) / (A + b + NR + NS)) * reg, Ea * ((Ea + A + NR) / (Ea - NR)) * reg
        elif spec.form == "eon":                                               # Eon...:291-293
            sA, sb, sE = A * (A / (A + NS + NR)), b * ((A + b + NR) / (A + b + NR + NS)), Ea * ((Ea + A + NS + NR) / (Ea - NS - NR))
        elif spec.form == "eoff":                                              # Eoff...:208-210
            sA, sb, sE = A * (A / (A + NS + NR)), b * ((A + b + NR) / (A + b + NR + NS)), Ea * ((Ea + A + b + NS + NR) / (Ea - b - NS - NR))
        else:
            raise ValueError(spec.form)
        self.slope_A, self.slope_b, self.slope_Ea = (s.to(device) for s in (sA, sb, sE))
        # projector onto the element-balance null space with the reference's 1e-4 ridge: N (N^T N + eps I)^-1 N^T
        N = self.E_null
        self.M = torch.linalg.solve(N.T @ N + 1e-4 * torch.eye(N.shape[1], dtype=f32, device=N.device), N.T)

    def __call__(self, p: torch.Tensor):
        s = self.spec
        w_b = torch.abs(p[:NR]) * self.slope_A
        w_in_b = p[NR:2 * NR] * self.slope_b
        w_in_Ea = torch.abs(p[2 * NR:3 * NR] * self.slope_Ea)
        w_out = p[3 * NR:(NS + 3) * NR].view(NS, NR)
        # column-wise  E_null @ solve(N^T N + eps I, N^T w_out[:, i])  (:207-213) as one product
        w_adj = self.E_null @ (self.M @ w_out)
        w_adj = torch.clamp(w_adj, s.wout[0], s.wout[1])
        w_in_only = torch.clamp(-w_adj, s.win[0], s.win[1])
        w_in_Ea = torch.clamp(w_in_Ea, s.Ea[0], s.Ea[1])
        w_in_b = torch.clamp(w_in_b, s.b[0], s.b[1])
        w_b = torch.clamp(w_b, s.A[0], s.A[1])
        w_in = torch.cat([w_in_only, w_in_Ea.unsqueeze(0), w_in_b.unsqueeze(0)], dim=0)
        return w_in, w_b, w_adj


@dataclass
class TrainingBatch:
    """Device-resident training conditions of one rank (knot-major SoA like the sweep)."""
    T0: torch.Tensor       # [n] float32
    c0: torch.Tensor       # [n] float32
    tgrid: torch.Tensor    # [801, n] float32
    Tprof: torch.Tensor | None   # [801, n] float32 (Eon trainer) or None (isothermal)
    ref: torch.Tensor      # [801, 7, n] float32 labels, mol/m3
    yscale: torch.Tensor   # [7, n] float32 = clamp(max_t - min_t, 1e-6, inf)  (:105)

    @property
    def n(self):
        return self.T0.numel()

    @staticmethod
    def yscale_from_labels(ref: torch.Tensor) -> torch.Tensor:
        return torch.clamp(ref.amax(dim=0) - ref.amin(dim=0), min=1e-6)

    def subset(self, idx) -> "TrainingBatch":
        """The conditions `idx` (any order) as a new contiguous batch -- a mini-batch or the train/valid/test part."""
        ix = torch.as_tensor(np.asarray(idx, np.int64), device=self.T0.device)
        pick = lambda x, d: None if x is None else x.index_select(d, ix).contiguous()
        return TrainingBatch(pick(self.T0, 0), pick(self.c0, 0), pick(self.tgrid, 1), pick(self.Tprof, 1), pick(self.ref, 2),
                             pick(self.yscale, 1))


def allreduce_packed(packed: torch.Tensor, group=None) -> torch.Tensor:
    """Sum [grad(189) | loss | count] over the ranks (the step's only collective); identity without a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


# (rtol, atol) of the free-stepping forward pass of the isothermal trainers ("dp54w"): knot states within ~1e-7 of the converged
# trajectory, i.e. at least as close as the knot-limited pass they replace at the reference's 1e-4 / 1e-6
FREE_STEP_TOLERANCE = (1.0e-7, 1.0e-10)


class CrnnTrainer:
    """loss / gradient / optimiser step for the flat parameter vector p[189].

    substeps: RK4 sub-steps of the adjoint sweep per knot interval.  One is the default: against float64 central differences of
    the converged CPU loss the gradient is then within 5.6e-6 of its scale (two: 1.2e-6, four: 1.2e-6 -- the finite differences' own
    floor; tests/test_training.py), and with parameters 5 % off their trained values within 2.7e-4 of the four-sub-step gradient
    (two: 8.6e-5; tools/adjoint_substeps.py, profiles/r02q_adjoint_substeps.jsonl) -- two orders below the solver-tolerance
    noise of the reference's own back-propagated gradient -- for 1.9 instead of 3.5 ms per 640 conditions."""

    def __init__(self, batch: TrainingBatch, spec: ConverterSpec = WIDE_LLNL, clamps=TRAINING_WIDE_CLAMPS, rtol=1e-4, atol=1e-6,
                 substeps=1, lr=5e-4, weight_decay=1e-4, clip=10.0, group=None, settings: "TrainerSettings | None" = None):
        if settings is not None:
            clamps, rtol, atol, lr, weight_decay, clip = (settings.clamps, settings.rtol, settings.atol, settings.lr,
                                                          settings.weight_decay, settings.clip)
        self.settings = settings
        self.batch, self.clamps, self.rtol, self.atol, self.substeps = batch, clamps, rtol, atol, substeps
        self.device = batch.T0.device
        self.converter = ParameterConverter(spec)
        self.clip, self.group = clip, group
        self.lr, self.weight_decay = lr, weight_decay
        self.opt = None
        # integrate() only needs a CRNN handle; reuse Surrogate's plumbing without MLPs
        self._sur = Surrogate.__new__(Surrogate)
        self._sur.device = self.device
        self._sur.energy_on = batch.Tprof is not None
        # forward integrator: every step ends on a knot of the label grid (dense output), i.e. the knot-limited regime in
        # which the explicit fast path is 2-4x cheaper than the Rosenbrock kernel (stiff conditions fall back to it)
        # ("bs23w": that integrator with one condition per warp -- a training batch is a few hundred conditions, so the pass is
        # latency-bound and nine lanes per right-hand side make it ~4x shorter than one thread per condition, "bs23")
        # Isothermal batches (the wide / narrow Eoff trainers): "dp54w" -- nothing ties a step to a knot at constant temperature, so
        # the pass takes free Dormand-Prince steps (a few dozen instead of 800) and reads the 801 knot states off the method's
        # continuous extension, as torchdiffeq's dopri5 does for the reference (WIDE_Eoff_surrogate_model_training.py:383).  A
        # free step's error sits AT the tolerance (a knot-limited one far below it), so the pass runs at FREE_STEP_TOLERANCE or the
        # trainer's own tolerances, whichever is tighter.
        self.forward_method = os.environ.get("PFR_TRAIN_FORWARD", "dp54w" if batch.Tprof is None else "bs23w")
        self._crnn, self._bufs, self.failed_last = None, {}, 0
        # gradient kernels: "staged" = three kernels with a workspace of 3.8 MB per condition (pfr_loss_grad_staged), "warp" = the
        # single kernel, one condition per warp, no workspace (pfr_loss_grad); "auto" = staged while the workspace stays under 8 GB
        self.adjoint = os.environ.get("PFR_TRAIN_ADJOINT", "auto")
        self._adj_ws = None

    # ---------------------------------------------------------------- device part
    def _model(self, w_in, w_b, w_out) -> CrnnModel:
        """ONE model handle for the trainer's lifetime: the parameters of a step are written into it (crnn_model_update; they reach
        the kernels by value at launch), instead of a handle created and destroyed per step."""
        params = CRNNParams(w_in, w_b, w_out)
        if self._crnn is None:
            self._crnn = CrnnModel(params, self.clamps)
        else:
            self._crnn.update(params)
        self._sur.crnn = self._crnn
        return self._crnn

    def forward(self, w_in, w_b, w_out, batch: TrainingBatch | None = None):
        """Raw knot states [801, 9, n] (float64) of the current parameters; status [n]."""
        crnn = self._model(w_in, w_b, w_out)
        b = batch or self.batch
        method = self.forward_method if b.Tprof is None or self.forward_method != "dp54w" else "bs23w"
        rtol, atol = self.rtol, self.atol
        if method == "dp54w":
            rtol, atol = min(rtol, FREE_STEP_TOLERANCE[0]), min(atol, FREE_STEP_TOLERANCE[1])
        res = self._sur.integrate(b.T0, b.c0, tgrid=b.tgrid, Tprof=b.Tprof, rtol=rtol, atol=atol, dense=True, dense_raw=True,
                                  method=method)
        return crnn, res

    def _buffers(self, n: int):
        """[grad(189) | loss] x n per-condition rows, the packed [grad | loss | count] vector and its page-locked host copy, kept
        per batch size so that a step allocates nothing but the forward pass's own outputs."""
        buf = self._bufs.get(n)
        if buf is None:
            buf = self._bufs[n] = (torch.empty((NPAR + 1, n), dtype=torch.float64, device=self.device),
                                   torch.empty(NPAR + 3, dtype=torch.float64, device=self.device),
                                   torch.empty(NPAR + 3, dtype=torch.float64).pin_memory())
        return buf

    def packed_loss_grad(self, w_in, w_b, w_out, batch: TrainingBatch | None = None) -> torch.Tensor:
        """This rank's [sum of gradients (189) | sum of losses | number of conditions summed | number of conditions attempted],
        float64 on the device (the vector the step all-reduces); the sums run over the conditions whose forward integration
        succeeded."""
        b = batch or self.batch
        crnn, res = self.forward(w_in, w_b, w_out, b)
        n = b.n
        rows, packed, _ = self._buffers(n)
        need = _lib.lib().pfr_loss_grad_workspace_bytes(n, self.substeps) if self.substeps > 0 else 0
        if need and (self.adjoint == "staged" or (self.adjoint == "auto" and need <= 8 << 30)):
            if self._adj_ws is None or self._adj_ws.numel() < need:
                self._adj_ws = None   # release the smaller one first
                self._adj_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            _lib.check(_lib.lib().pfr_loss_grad_staged(crnn.handle, n, _ptr(b.T0), _ptr(b.tgrid), _ptr(b.Tprof), _ptr(res.dense), _ptr(b.ref),
                                                       _ptr(b.yscale), self.substeps, _ptr(rows[NPAR]), _ptr(rows), _ptr(self._adj_ws),
                                                       self._adj_ws.numel(), _stream()), "pfr_loss_grad_staged")
        else:
            _lib.check(_lib.lib().pfr_loss_grad(crnn.handle, n, _ptr(b.T0), _ptr(b.tgrid), _ptr(b.Tprof), _ptr(res.dense), _ptr(b.ref),
                                                _ptr(b.yscale), self.substeps, _ptr(rows[NPAR]), _ptr(rows), _stream()), "pfr_loss_grad")
        # a failed trajectory (status != 0) enters neither the sums nor the count
        _lib.check(_lib.lib().pfr_reduce_rows_ok(_ptr(rows), NPAR + 1, n, _ptr(res.status), _ptr(packed), _stream()), "pfr_reduce_rows_ok")
        packed[NPAR + 2:].fill_(float(n))
        return packed

    def loss_grad_w(self, w_in, w_b, w_out, batch: TrainingBatch | None = None):
        """(sum of per-condition losses, sum of per-condition gradients [189], failed count) on this rank, float64 CUDA."""
        b = batch or self.batch
        packed = self.packed_loss_grad(w_in, w_b, w_out, b).clone()
        return packed[NPAR], packed[:NPAR], int(round(float(packed[NPAR + 2] - packed[NPAR + 1])))

    def _reduce_to_host(self, packed: torch.Tensor, n: int) -> torch.Tensor:
        """All-reduce the packed vector over the ranks and bring it to the host: one collective, one page-locked copy, one wait."""
        packed = allreduce_packed(packed, self.group)
        host = self._buffers(n)[2]
        host.copy_(packed, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host

    # ---------------------------------------------------------------- host part (189 numbers)
    def loss_and_grad(self, p: torch.Tensor, idx=None):
        """Mean loss over all ranks' (successfully integrated) conditions -- or over this rank's conditions `idx`, a mini-batch --
        and its gradient with respect to p (float32 CPU tensors); third value: failed trajectories (all ranks), which entered
        neither the mean nor the gradient."""
        b = self.batch if idx is None else self.batch.subset(idx)
        pc = p.detach().to("cpu", torch.float32).requires_grad_(True)
        w_in, w_b, w_out = self.converter(pc)
        packed = self._reduce_to_host(self.packed_loss_grad(w_in.detach().numpy(), w_b.detach().numpy(), w_out.detach().numpy(), b), b.n)
        count = max(float(packed[NPAR + 1]), 1.0)
        bad = int(round(float(packed[NPAR + 2] - packed[NPAR + 1])))   # failed trajectories, all ranks
        g = (packed[:NPAR] / count).to(torch.float32)
        g_in, g_b, g_out = g[:99].view(11, 9), g[99:108], g[108:].view(9, 9)
        (gp,) = torch.autograd.grad((w_in, w_b, w_out), pc, (g_in, g_b, g_out))
        self.failed_last = bad
        return float(packed[NPAR]) / count, gp, bad

    def loss(self, p: torch.Tensor, idx=None) -> float:
        """Mean loss over all ranks' conditions without a parameter gradient (the validation / test loop, :424-431)."""
        b = self.batch if idx is None else self.batch.subset(idx)
        with torch.no_grad():
            w_in, w_b, w_out = self.converter(p.detach().to("cpu", torch.float32))
        packed = self._reduce_to_host(self.packed_loss_grad(w_in.numpy(), w_b.numpy(), w_out.numpy(), b), b.n)
        return float(packed[NPAR]) / max(float(packed[NPAR + 1]), 1.0)

    def step(self, p: torch.Tensor, idx=None):
        """One optimiser step on p (a CPU float32 leaf tensor): gradient over the whole batch (or the mini-batch `idx`),
        clip_grad_norm_(10), AdamW(5e-4, wd 1e-4)."""
        if self.opt is None:
            self.opt = torch.optim.AdamW([p], lr=self.lr, weight_decay=self.weight_decay)
        loss, gp, bad = self.loss_and_grad(p, idx)
        self.opt.zero_grad()
        p.grad = gp
        if self.clip:
            torch.nn.utils.clip_grad_norm_([p], self.clip)
        self.opt.step()
        return loss, bad


def save_history(path: str, history: dict, final: tuple | None = None, p: torch.Tensor | None = None) -> None:
    """np.savez(save_path, **history) as the reference writes it (WIDE_Eoff...:445-471): keys train_loss, valid_loss,
    parameters (object array of {'w_in','w_b','w_out'} float32 dicts) and, at the end of training, final_parameters and
    updated_p.  The files load with the reference's load_npz_parameters (parameters[-1])."""
    data = dict(history)
    if final is not None:
        data["final_parameters"] = {"w_in": final[0], "w_b": final[1], "w_out": final[2]}
    if p is not None:
        data["updated_p"] = p.detach().cpu().numpy()
    np.savez(path, **data)


def train(trainer: "CrnnTrainer", p: torch.Tensor, epochs: int, valid: "CrnnTrainer | None" = None, save_path: str | None = None,
          steps_per_epoch: int = 1, log=None, batch_size: int | None = None, shuffle_seed: int | None = 0,
          lr_factor: float | None = None) -> dict:
    """Trainer.train (WIDE_Eoff...:398-476; Eon...:498-566).  Per epoch: the optimiser steps, the validation loss,
    ReduceLROnPlateau(mode min, patience 5, threshold 1e-4 rel; factor 0.8 wide :501, 0.6 Eoff :515, 0.5 Eon :598), the
    history append and the per-epoch np.savez.  Returns the history dict.

    batch_size None : `steps_per_epoch` steps, each on the rank's whole training shard (the fast schedule).
    batch_size B    : one pass over the shuffled conditions in mini-batches of B; B = 1 is the reference's own schedule
                      (random.shuffle(train_idx), one optimiser step per sample, train loss = mean of the per-step losses).
    With several ranks every rank must hold the same number of conditions so that the step counts agree."""
    history = {"train_loss": [], "valid_loss": [], "parameters": []}
    failed_total = 0
    sched = None
    if lr_factor is None:
        lr_factor = trainer.settings.lr_factor if trainer.settings is not None else 0.8
    rng = np.random.default_rng(shuffle_seed)
    for epoch in range(epochs):
        if batch_size is None:
            out = [trainer.step(p) for _ in range(steps_per_epoch)]
            losses, failed_total = [o[0] for o in out], failed_total + sum(o[1] for o in out)
        else:
            order = np.arange(trainer.batch.n) if shuffle_seed is None else rng.permutation(trainer.batch.n)
            out = [trainer.step(p, order[s:s + batch_size]) for s in range(0, len(order), batch_size)]
            losses, failed_total = [o[0] for o in out], failed_total + sum(o[1] for o in out)
        if sched is None:
            sched = torch.optim.lr_scheduler.ReduceLROnPlateau(trainer.opt, mode="min", factor=lr_factor, patience=5, threshold=1e-4,
                                                               threshold_mode="rel")
        history["train_loss"].append(float(np.mean(losses)))
        vloss = (valid or trainer).loss(p)
        sched.step(vloss)
        history["valid_loss"].append(float(vloss))
        w_in, w_b, w_out = (x.detach().cpu().numpy() for x in trainer.converter(p.detach()))
        history["parameters"].append({"w_in": w_in, "w_b": w_b, "w_out": w_out})
        if failed_total and log:
            log(epoch, f"{failed_total} forward trajectories failed so far; they were left out of the loss and the gradient", vloss, None)
        if log:
            log(epoch, history["train_loss"][-1], vloss, trainer.opt.param_groups[0]["lr"])
        if save_path:
            save_history(save_path, history)
    if save_path:
        w = tuple(x.detach().cpu().numpy() for x in trainer.converter(p.detach()))
        save_history(save_path, history, final=w, p=p)
    history["failed_trajectories"] = failed_total   # (after the last save: the reference's history files hold the three keys above only)
    return history


def synthetic_labels(sur: Surrogate, teacher: CRNNParams, T, P, clamps=TRAINING_WIDE_CLAMPS, rtol=1e-10, atol=1e-12) -> TrainingBatch:
    """Training batch with teacher-generated labels (the reference's Cantera label files are not shipped): time grid
    from the surrogate's time MLP at (T, P, 1.0 m, 2.5 m/s), labels = teacher CRNN trajectories at the knots.  With an
    Eoff model set the batch is isothermal; with an Eon set the temperature profile comes from the temperature MLP at
    (T, P), which is what the Eon trainer integrates along (Eon_surrogate_model_training.py:118-180)."""
    T = torch.as_tensor(np.asarray(T, np.float32)).to(sur.device)
    P = torch.as_tensor(np.asarray(P, np.float32)).to(sur.device)
    c0 = sur.inlet_concentration(T, P)
    tgrid, _ = sur.time_grid(T, P, None, None)
    Tprof = sur.temp_profile(T, P) if sur.energy_on else None
    helper = Surrogate.__new__(Surrogate)
    helper.device = sur.device
    helper.energy_on = sur.energy_on
    helper.crnn = CrnnModel(teacher, clamps)
    res = helper.integrate(T, c0, tgrid=tgrid, Tprof=Tprof, rtol=rtol, atol=atol, dense=True).raise_on_failure()
    ref = res.dense[:, :7, :].to(torch.float32).contiguous()
    return TrainingBatch(T, c0, tgrid, Tprof, ref, TrainingBatch.yscale_from_labels(ref).contiguous())
