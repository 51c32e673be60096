"""A few optimisation steps of the device MLP trainer on random data, for ncu launch lists: python tools/mlp_train_probe.py [steps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.mlp_training import MlpTrainer, TEMP_2D_SETTINGS, initial_parameters
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
w0, b0 = initial_parameters(2, seed=0)
tr = MlpTrainer(w0, b0, TEMP_2D_SETTINGS)
x, y = torch.rand((32, 2), device="cuda"), torch.rand((32, 800), device="cuda")
loss = torch.zeros((), device="cuda")
for _ in range(steps):
    tr.step(x, y, 1e-3, loss)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    tr.step(x, y, 1e-3, loss)
e1.record(); torch.cuda.synchronize()
print("loss", float(loss), "us/step", e0.elapsed_time(e1) * 1e3 / 200)
