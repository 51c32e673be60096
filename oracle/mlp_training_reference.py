"""TEST INFRASTRUCTURE (oracle): torch-CPU restatement of the predictor-MLP training scripts
  TEMP_PRED_MODEL_TRAINING/temp_profile_model_training_2D.py   and   TIME_PRED_MODEL_TRAINING/time_profile_model_training_4D.py
for SURVEY 8(f) item 4.  Unlike the CRNN path (torchdiffeq), everything these scripts call is plain torch, which IS
installed here: this file runs the reference's own operators (nn.Linear, nn.ReLU, nn.MSELoss, torch.optim.Adam,
lr_scheduler.StepLR) on the CPU, so the parity of the device trainer is pinned against the real framework, not against a
re-derivation.  Only tests/ and bench.py's CPU leg may import it."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


class MultiLayerPerceptron(nn.Module):
    """…2D.py:104-121 / …4D.py:137-158 (input_node 2 or 4, neurons 512, output_node 800)."""

    def __init__(self, input_node: int, output_node: int = 800, neurons: int = 512):
        super().__init__()
        self.fc1 = nn.Linear(input_node, neurons)
        self.relu1 = nn.ReLU()
        self.fc2 = nn.Linear(neurons, neurons)
        self.relu2 = nn.ReLU()
        self.fc3 = nn.Linear(neurons, neurons)
        self.relu3 = nn.ReLU()
        self.fc4 = nn.Linear(neurons, output_node)

    def forward(self, x):
        x = self.relu1(self.fc1(x))
        x = self.relu2(self.fc2(x))
        x = self.relu3(self.fc3(x))
        return self.fc4(x)


def make_model(weights, biases) -> MultiLayerPerceptron:
    m = MultiLayerPerceptron(int(np.asarray(weights[0]).shape[1]))
    with torch.no_grad():
        for i, fc in enumerate((m.fc1, m.fc2, m.fc3, m.fc4)):
            fc.weight.copy_(torch.as_tensor(np.asarray(weights[i], np.float32)))
            fc.bias.copy_(torch.as_tensor(np.asarray(biases[i], np.float32)))
    return m


def run_steps(weights, biases, batches, lrs, betas=(0.9, 0.999), eps=1e-8):
    """The inner loop of …2D.py:145-160 on given mini-batches [(x, y), …] with the learning rate of each step given (StepLR
    on the caller's side).  Returns (per-step losses, final weights, final biases)."""
    torch.set_num_threads(max(1, torch.get_num_threads()))
    model = make_model(weights, biases)
    model.train()
    criterion = nn.MSELoss()
    opt = torch.optim.Adam(model.parameters(), lr=lrs[0], betas=betas, eps=eps)
    losses = []
    for (x, y), lr in zip(batches, lrs):
        for gparam in opt.param_groups:
            gparam["lr"] = lr
        out = model(torch.as_tensor(x))
        loss = criterion(out, torch.as_tensor(y))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    fcs = (model.fc1, model.fc2, model.fc3, model.fc4)
    return np.asarray(losses), [fc.weight.detach().numpy().copy() for fc in fcs], [fc.bias.detach().numpy().copy() for fc in fcs]


def train(parts, settings, weights, biases, batch_index_lists, num_epochs):
    """The epoch loop of …2D.py:137-177 with externally supplied shuffles (batch_index_lists[epoch] = (train batches, valid
    batches) as index arrays), so that the device trainer can be fed the identical sequence.  Returns (history_train,
    history_valid, model) with the scripts' running_loss bookkeeping."""
    from torch.optim import lr_scheduler
    model = make_model(weights, biases)
    criterion = nn.MSELoss()
    opt = torch.optim.Adam(model.parameters(), lr=settings.learning_rate, betas=settings.betas, eps=settings.eps)
    sched = lr_scheduler.StepLR(opt, step_size=settings.lr_step_size, gamma=settings.lr_gamma)
    xt, yt = (torch.as_tensor(a) for a in parts["training"])
    xv, yv = (torch.as_tensor(a) for a in parts["valid"])
    hist_t, hist_v = [], []
    for epoch in range(num_epochs):
        tb, vb = batch_index_lists[epoch]
        model.train()
        running = 0.0
        for idx in tb:
            loss = criterion(model(xt[idx]), yt[idx])
            opt.zero_grad()
            loss.backward()
            opt.step()
            running += loss.item()
        hist_t.append(running / len(tb))
        sched.step()
        model.eval()
        with torch.no_grad():
            for idx in vb:
                running += criterion(model(xv[idx]), yv[idx]).item()
        hist_v.append(running / len(tb))
    return hist_t, hist_v, model
