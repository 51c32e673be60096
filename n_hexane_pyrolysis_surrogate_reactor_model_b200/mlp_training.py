"""Device training of the predictor MLPs: host mirror of
  TEMP_PRED_MODEL_TRAINING/temp_profile_model_training_2D.py   (T, P)        -> 800 temperatures
  TIME_PRED_MODEL_TRAINING/time_profile_model_training_4D.py   (T, P, L, u0) -> 800 residence times
Same data scaling and 80/10/10 split (train_test_split, random_state 2024), same network, loss, optimiser, StepLR and
epoch loop; every optimisation step is pfr_mlp_trainer_step (csrc/mlp_train.cuh).  Writes the `.pth` state dict and the
`.pkl` min/max file the reference's inference scripts (and containers.load_mlp) read.  The Cantera label files of the
reference are not shipped, so tests and the bench train against labels produced by a shipped (teacher) MLP."""
from __future__ import annotations

import ctypes
import pickle
from collections import OrderedDict
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from .containers import HIDDEN, MLPParams, NTOTAL

NOUT = NTOTAL - 1
INPUT_SCALE = {2: np.asarray([[870.0, 1.0], [1150.0, 3.0]]),                       # temp_profile_model_training_2D.py:46-47
               4: np.asarray([[870.0, 1.0, 0.5, 2.5], [1150.0, 3.0, 1.0, 5.0]])}   # time_profile_model_training_4D.py:44-45


@dataclass
class MlpTrainSettings:
    """Constants of the two scripts (…2D.py:22-26,127-128; …4D.py:22-26,164-165)."""
    in_dim: int = 2
    batch_size: int = 32
    learning_rate: float = 1.0e-3
    num_epochs: int = 20            # 3000 in the time-profile script
    lr_step_size: int = 100         # StepLR(step_size=100, gamma=0.6), stepped once per epoch
    lr_gamma: float = 0.6
    betas: tuple = (0.9, 0.999)     # torch.optim.Adam defaults
    eps: float = 1.0e-8
    split_seed: int = 2024
    shuffle_seed: int = 0           # the scripts do not seed their DataLoader; a seed makes a run repeatable


TEMP_2D_SETTINGS = MlpTrainSettings(in_dim=2, num_epochs=20)
TIME_4D_SETTINGS = MlpTrainSettings(in_dim=4, num_epochs=3000)


@dataclass
class MlpDataset:
    """TemperatureDataset / TimeDataset: inputs in physical units ([K, bar] or [K, bar, m, m/s]), outputs [n, 800] in physical
    units (K or s); scaled to [0, 1] by the fixed input ranges and by the global min / max of the outputs, then split
    80 / 10 / 10 with two train_test_split calls (random_state 2024), exactly as the scripts do."""
    inputs: np.ndarray
    outputs: np.ndarray
    split_seed: int = 2024
    parts: dict = field(init=False)
    output_scale: np.ndarray = field(init=False)

    def __post_init__(self):
        from sklearn.model_selection import train_test_split
        x = np.array(self.inputs, dtype=float)
        y = np.array(self.outputs, dtype=float)
        sc = INPUT_SCALE[x.shape[1]]
        self.output_scale = np.asarray([np.min(y), np.max(y)])
        x = (x - sc[0]) / (sc[1] - sc[0])
        y = (y - self.output_scale[0]) / (self.output_scale[1] - self.output_scale[0])
        x_tr, x_te, y_tr, y_te = train_test_split(x, y, test_size=0.2, random_state=self.split_seed)
        x_va, x_te, y_va, y_te = train_test_split(x_te, y_te, test_size=0.5, random_state=self.split_seed)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)   # Dataset.__getitem__: torch.from_numpy(...).float()
        self.parts = {"training": (f(x_tr), f(y_tr)), "valid": (f(x_va), f(y_va)), "test": (f(x_te), f(y_te))}


def initial_parameters(in_dim: int, seed: int = 0):
    """nn.Linear's default initialisation of fc1..fc4 (what MultiLayerPerceptron(output_node) starts from), seeded."""
    g = torch.Generator().manual_seed(seed)
    w, b = [], []
    for fan_in, fan_out in ((in_dim, HIDDEN), (HIDDEN, HIDDEN), (HIDDEN, HIDDEN), (HIDDEN, NOUT)):
        bound = 1.0 / np.sqrt(fan_in)   # kaiming_uniform_(a=sqrt(5)) on [out, in] and the bias bound are both 1/sqrt(fan_in)
        w.append(((torch.rand((fan_out, fan_in), generator=g) * 2 - 1) * bound).numpy())
        b.append(((torch.rand(fan_out, generator=g) * 2 - 1) * bound).numpy())
    return w, b


class MlpTrainer:
    """Parameters, Adam moments and activations of one predictor MLP on the GPU (pfr_mlp_trainer_t)."""

    def __init__(self, weights, biases, settings: MlpTrainSettings, device="cuda"):
        if not torch.cuda.is_available():
            raise _lib.PfrError("no CUDA device: this package has no CPU path")
        self.settings = settings
        self.device = torch.device(device)
        self.in_dim = int(np.asarray(weights[0]).shape[1])
        self._w = [np.ascontiguousarray(a, dtype=np.float32) for a in weights]
        self._b = [np.ascontiguousarray(a, dtype=np.float32) for a in biases]
        fp = lambda arrs: (_lib.c_float_p * 4)(*[a.ctypes.data_as(_lib.c_float_p) for a in arrs])
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().pfr_mlp_trainer_create(self.in_dim, fp(self._w), fp(self._b), ctypes.byref(h)), "pfr_mlp_trainer_create")
        self.handle = h
        self.steps = 0

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().pfr_mlp_trainer_destroy(self.handle)
        except Exception:
            pass

    @staticmethod
    def _ptr(t):
        return ctypes.c_void_p(t.data_ptr())

    def step(self, x: torch.Tensor, y: torch.Tensor, lr: float, loss_out: torch.Tensor | None = None) -> torch.Tensor:
        """One optimisation step on a mini-batch (x [B, in], y [B, 800], float32 on the device, B <= 32); returns the
        0-d device tensor holding the batch loss before the update."""
        loss = torch.empty((), dtype=torch.float32, device=self.device) if loss_out is None else loss_out
        s = self.settings
        _lib.check(_lib.lib().pfr_mlp_trainer_step(self.handle, self._ptr(x), self._ptr(y), x.shape[0], float(lr), s.betas[0], s.betas[1],
                                                   s.eps, self._ptr(loss), torch.cuda.current_stream().cuda_stream), "pfr_mlp_trainer_step")
        self.steps += 1
        return loss

    def loss(self, x: torch.Tensor, y: torch.Tensor, loss_out: torch.Tensor | None = None) -> torch.Tensor:
        loss = torch.empty((), dtype=torch.float32, device=self.device) if loss_out is None else loss_out
        _lib.check(_lib.lib().pfr_mlp_trainer_loss(self.handle, self._ptr(x), self._ptr(y), x.shape[0], self._ptr(loss),
                                                   torch.cuda.current_stream().cuda_stream), "pfr_mlp_trainer_loss")
        return loss

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(self.device, torch.float32).contiguous()
        out = torch.empty((x.shape[0], NOUT), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().pfr_mlp_trainer_forward(self.handle, self._ptr(x), x.shape[0], self._ptr(out),
                                                      torch.cuda.current_stream().cuda_stream), "pfr_mlp_trainer_forward")
        return out

    def parameters(self):
        w = [np.empty_like(a) for a in self._w]
        b = [np.empty_like(a) for a in self._b]
        fp = lambda arrs: (_lib.c_float_p * 4)(*[a.ctypes.data_as(_lib.c_float_p) for a in arrs])
        _lib.check(_lib.lib().pfr_mlp_trainer_read(self.handle, fp(w), fp(b)), "pfr_mlp_trainer_read")
        return w, b

    def state_dict(self) -> OrderedDict:
        w, b = self.parameters()
        sd = OrderedDict()
        for i in range(4):
            sd[f"fc{i + 1}.weight"] = torch.from_numpy(w[i].copy())
            sd[f"fc{i + 1}.bias"] = torch.from_numpy(b[i].copy())
        return sd

    def save(self, pth_path: str, pkl_path: str, output_scale) -> None:
        """torch.save(model.state_dict(), …pth) and the {'min', 'max'} pickle (…2D.py:63-65,179; …4D.py:72-74,223)."""
        torch.save(self.state_dict(), pth_path)
        with open(pkl_path, "wb") as f:
            pickle.dump({"min": np.float64(output_scale[0]), "max": np.float64(output_scale[1])}, f)

    def mlp_params(self, output_scale) -> MLPParams:
        w, b = self.parameters()
        return MLPParams(w, b, float(output_scale[0]), float(output_scale[1]))


def step_lr(settings: MlpTrainSettings, epoch: int) -> float:
    """Learning rate in force during epoch `epoch` (0-based) under lr_scheduler.StepLR(step_size, gamma) stepped once per epoch
    after the training batches (…2D.py:128,162; …4D.py:165,192)."""
    return settings.learning_rate * settings.lr_gamma ** (epoch // settings.lr_step_size)


def epoch_batches(n: int, batch_size: int, generator: torch.Generator):
    """DataLoader(shuffle=True, drop_last=False): a fresh permutation per epoch, consecutive slices of batch_size."""
    perm = torch.randperm(n, generator=generator)
    return [perm[i:i + batch_size] for i in range(0, n, batch_size)]


def train(dataset: MlpDataset, settings: MlpTrainSettings, weights=None, biases=None, init_seed: int = 0, num_epochs: int | None = None,
          device="cuda", on_epoch=None):
    """The scripts' training loop (…2D.py:137-177, …4D.py:166-210).  Returns (trainer, history_train, history_valid).
    As in the scripts, history_train[e] is the mean batch loss of the epoch, and history_valid[e] is (sum of the training
    batch losses + sum of the validation batch losses) / number of training batches -- running_loss is not reset before
    the validation loop there."""
    if weights is None:
        weights, biases = initial_parameters(settings.in_dim, init_seed)
    tr = MlpTrainer(weights, biases, settings, device)
    dev = tr.device
    xt, yt = (torch.from_numpy(a).to(dev) for a in dataset.parts["training"])
    xv, yv = (torch.from_numpy(a).to(dev) for a in dataset.parts["valid"])
    g = torch.Generator().manual_seed(settings.shuffle_seed)
    nb_train = (xt.shape[0] + settings.batch_size - 1) // settings.batch_size
    nb_valid = (xv.shape[0] + settings.batch_size - 1) // settings.batch_size
    losses = torch.zeros(nb_train + nb_valid, dtype=torch.float32, device=dev)
    hist_t, hist_v = [], []
    for epoch in range(settings.num_epochs if num_epochs is None else num_epochs):
        lr = step_lr(settings, epoch)
        for j, idx in enumerate(epoch_batches(xt.shape[0], settings.batch_size, g)):
            idx = idx.to(dev)
            tr.step(xt[idx].contiguous(), yt[idx].contiguous(), lr, losses[j])
        for j, idx in enumerate(epoch_batches(xv.shape[0], settings.batch_size, g)):
            idx = idx.to(dev)
            tr.loss(xv[idx].contiguous(), yv[idx].contiguous(), losses[nb_train + j])
        host = losses.double().cpu().numpy()             # one synchronisation per epoch instead of one loss.item() per batch
        run = float(host[:nb_train].sum())
        hist_t.append(run / nb_train)
        hist_v.append((run + float(host[nb_train:].sum())) / nb_train)
        if on_epoch is not None:
            on_epoch(epoch, hist_t[-1], hist_v[-1], lr)
    return tr, hist_t, hist_v


def evaluate_test_set(trainer: MlpTrainer, dataset: MlpDataset) -> dict:
    """The scripts' test-set figures (…2D.py:181-215 and the blocks after it): mean accuracy 100 (1 - MAPE), R^2, per-case RMSE /
    MAE / relative error in physical units."""
    x, y = dataset.parts["test"]
    lo, hi = dataset.output_scale
    pred = trainer.forward(torch.from_numpy(x)).double().cpu().numpy() * (hi - lo) + lo
    true = y.astype(np.float64) * (hi - lo) + lo
    acc = (1 - np.abs(pred - true) / np.abs(true)) * 100
    ss_res, ss_tot = np.sum((true - pred) ** 2), np.sum((true - true.mean()) ** 2)
    rmse = np.sqrt(np.mean((true - pred) ** 2, axis=1))
    mae = np.mean(np.abs(true - pred), axis=1)
    rel = np.mean(np.abs(pred - true) / (np.abs(true) + 1e-12), axis=1) * 100
    return {"accuracy_mean": float(acc.mean()), "r2": float(1 - ss_res / ss_tot), "rmse_mean": float(rmse.mean()), "rmse_std": float(rmse.std()),
            "mae_mean": float(mae.mean()), "mae_std": float(mae.std()), "rel_error_mean": float(rel.mean()), "rel_error_std": float(rel.std())}
