"""Gradient of the training loss for 1, 2 and 4 RK4 sub-steps of the adjoint sweep per knot interval (640 wide-2D conditions):
time and deviation from the 4-sub-step gradient.  python tools/adjoint_substeps.py"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import CrnnTrainer, synthetic_labels

gold = os.path.join(ROOT, "tests", "golden")
a = np.load(os.path.join(gold, "conditions.npz"))["training_wide_2D"][:640]
sur = Surrogate(ModelSet.from_packed(os.path.join(gold, "containers", "LLNL.npz"), "Eoff"))
teacher = ModelSet.from_packed(os.path.join(gold, "containers", "LLNL.npz"), "Eoff", "Eoff_wide").crnn
batch = synthetic_labels(sur, teacher, a[:, 0].astype(np.float32), (a[:, 1] * 1e5).astype(np.float32))
kat = np.load(os.path.join(gold, "converter_kat.npz"))
p = torch.tensor(kat["LLNL_Eoff_wide/updated_p"]) + 0.05 * torch.randn(189, generator=torch.Generator().manual_seed(0))
out = {}
for sub in (4, 2, 1):
    tr = CrnnTrainer(batch, substeps=sub)
    w = [x.detach().numpy() for x in tr.converter(p)]
    tr.loss_grad_w(*w); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        l, g, bad = tr.loss_grad_w(*w)
    e1.record(); torch.cuda.synchronize()
    out[sub] = (float(l), g.cpu().numpy(), e0.elapsed_time(e1) / 5)
ref = out[4][1]
for sub in (4, 2, 1):
    l, g, ms = out[sub]
    print(json.dumps({"substeps": sub, "forward_plus_adjoint_ms": ms, "loss": l, "grad_dev_vs_4_substeps": float(np.max(np.abs(g - ref)) / np.max(np.abs(ref)))}))
