"""Training step vs the tolerance of the free-stepping forward pass (training.FREE_STEP_TOLERANCE): step / forward time, attempts, and the
error of knot states, loss and gradient against a knot-limited BS23 pass at 1e-10 / 1e-13 (640 wide-2D conditions, parameters 5 % off
their trained values).  python tools/r02_train_forward_tol.py  ->  profiles/r02w_train_forward_tolerance.jsonl"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200 import training as TR
gold = os.path.join(ROOT, "tests", "golden")
a = np.load(os.path.join(gold, "conditions.npz"))["training_wide_2D"][:640]
sur = Surrogate(ModelSet.from_packed(os.path.join(gold, "containers", "LLNL.npz"), "Eoff"))
teacher = ModelSet.from_packed(os.path.join(gold, "containers", "LLNL.npz"), "Eoff", "Eoff_wide").crnn
batch = TR.synthetic_labels(sur, teacher, a[:, 0].astype(np.float32), (a[:, 1] * 1e5).astype(np.float32))
kat = np.load(os.path.join(gold, "converter_kat.npz"))
p0 = torch.tensor(kat["LLNL_Eoff_wide/updated_p"]) + 0.05 * torch.randn(189, generator=torch.Generator().manual_seed(0))
tr = TR.CrnnTrainer(batch)
w = [x.detach().numpy() for x in tr.converter(p0)]
tr.forward_method = "bs23w"; tr.rtol, tr.atol = 1e-10, 1e-13
_, ref = tr.forward(*w); refd = ref.dense.clone(); lref, gref, _ = tr.loss_grad_w(*w); gref = gref.clone(); lref = float(lref)
tr.rtol, tr.atol = 1e-4, 1e-6
for tol in ((1e-7, 1e-10), (1e-7, 1e-9), (3e-7, 1e-9), (1e-6, 1e-9), (1e-6, 1e-8), (1e-5, 1e-8)):
    TR.FREE_STEP_TOLERANCE = tol
    tr.forward_method = "dp54w"
    _, res = tr.forward(*w)
    err = float(((res.dense - refd).abs() / refd.abs().clamp(min=1e-3)).max())
    l, g, _ = tr.loss_grad_w(*w)
    gerr = float((g - gref).abs().max() / gref.abs().max())
    p = p0.clone().requires_grad_(True)
    for _ in range(3): tr.step(p)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(10): tr.step(p)
    torch.cuda.synchronize(); dt = (time.time() - t0) / 10
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record(); tr.forward(*w); e[1].record(); torch.cuda.synchronize()
    st = res.stats.double()
    print(json.dumps(dict(tol=tol, step_ms=dt * 1e3, forward_ms=e[0].elapsed_time(e[1]), attempts_mean=float((st[0] + st[1]).mean()), attempts_max=float((st[0] + st[1]).max()),
                          knot_err_max=err, loss_rel_err=abs(float(l) - lref) / lref, grad_err=gerr)), flush=True)
    tr.opt = None
