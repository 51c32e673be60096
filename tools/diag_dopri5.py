import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from oracle import c_oracle as CO
g = np.load(os.path.join(ROOT, "tests/golden/reference_vectors.npz"))
ms = ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eoff")
s = Surrogate(ms)
tgd = torch.as_tensor(g["Eoff/tgrid"].T.copy()).cuda()
res = s.integrate(g["T"], g["c0"][:, 6], tgrid=tgd, method="dopri5", precision=32, dense=True)
st = res.stats.cpu().numpy()
print("gpu acc", st[0]); print("ref acc", g["Eoff/dopri5_stats"][:, 1])
print("gpu rej", st[1]); print("ref rej", g["Eoff/dopri5_stats"][:, 2])
print("gpu nfe", st[2]); print("ref nfe", g["Eoff/dopri5_stats"][:, 0])
y = res.y.cpu().numpy().T
ref = g["Eoff/dopri5_f32"][:, :, 800]
print("rel", (np.abs(y - ref) / np.maximum(np.abs(ref), 1e-3)).max(axis=1))
# bigger comparison vs the C oracle on 400 conditions
c = np.load(os.path.join(ROOT, "tests/golden/conditions.npz"))["independent_4D"]
T, P, L, U = c[:, 0].astype(np.float32), (c[:, 1] * 1e5).astype(np.float32), c[:, 2].astype(np.float32), c[:, 3].astype(np.float32)
grid, _ = s.time_grid(T, P, L, U)
c0 = s.inlet_concentration(T, P)
r2 = s.integrate(T, c0, tgrid=grid, method="dopri5", precision=32)
tg = grid.cpu().numpy().T.copy()
c0n = np.zeros((400, 9), np.float32); c0n[:, 6] = c0.cpu().numpy()
yo, _, so = CO.dopri5_batch(tg, np.repeat(T[:, None], 801, 1), c0n, ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out, nthreads=8)
sg = r2.stats.cpu().numpy()
same = (sg[0] == so[:, 1]) & (sg[1] == so[:, 2])
print("same step counts", same.mean(), "rel max", (np.abs(r2.y.cpu().numpy().T - yo) / np.maximum(np.abs(yo), 1e-3)).max(), "rel max where same", (np.abs(r2.y.cpu().numpy().T - yo) / np.maximum(np.abs(yo), 1e-3))[same].max())
r3 = s.integrate(T, c0, tgrid=grid, method="dopri5", precision=64)
yo64, _, so64 = CO.dopri5_batch(tg, np.repeat(T[:, None], 801, 1), c0n, ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out, precision=64, nthreads=8)
sg = r3.stats.cpu().numpy(); same = (sg[0] == so64[:, 1]) & (sg[1] == so64[:, 2])
print("f64 same step counts", same.mean(), "rel max", (np.abs(r3.y.cpu().numpy().T - yo64) / np.maximum(np.abs(yo64), 1e-3)).max())
