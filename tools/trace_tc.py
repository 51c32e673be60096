"""Per-tile timeline of the persistent tensor-core GEMM (development tool).

Builds a second copy of the library with -DPFR_TC_TRACE (time stamps from %globaltimer inside mlp_tc_gemm_kernel), runs
one 65536-condition chunk of the time MLP and prints, for the tiles one SM processed, when the MMA warp began / was
allowed to start (TMEM handed back) / committed, and when the epilogue saw the accumulators, finished reading TMEM and
finished storing.      python tools/trace_tc.py [layer 0|1|2]
"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200 import build as B
so = os.path.join(ROOT, "n_hexane_pyrolysis_surrogate_reactor_model_b200", "libcrnn_pfr_b200_trace.so")
if not os.path.exists(so):
    B.build(force=True, defines=["PFR_TC_TRACE"], out=so)
os.environ["CRNN_PFR_LIB"] = so
from n_hexane_pyrolysis_surrogate_reactor_model_b200 import _lib
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
if not torch.cuda.is_available():
    print("built", so); sys.exit(0)
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = 65536
ms = ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eoff")
T, P, L, U = [torch.as_tensor(np.asarray(v, np.float32)).cuda() for v in lhs_conditions(n, seed=1)]
s = Surrogate(ms)
s.time_grid(T, P, L, U, want_grid=False, want_end=True); torch.cuda.synchronize()
buf = torch.zeros((3, 7 * 512, 8), dtype=torch.int64, device="cuda")
lib = _lib.lib(); lib.pfr_dev_set_tc_trace.argtypes = [ctypes.c_void_p] * 3
lib.pfr_dev_set_tc_trace(buf[0].data_ptr(), buf[1].data_ptr(), buf[2].data_ptr())
s.time_grid(T, P, L, U, want_grid=False, want_end=True); torch.cuda.synchronize()
lib.pfr_dev_set_tc_trace(None, None, None)
b = buf[layer].cpu().numpy().astype(np.float64)
b = b[b[:, 0] > 0]
t0 = b[:, 0].min()
sel = np.where(b[:, 7] == b[5, 7])[0]
sel = sel[np.argsort(b[sel, 0])]
print("MMA thread: tile_begin, tmem_handed_back, committed | epilogue: bias_ready, accumulators_full, last_tmem_load, stored   (ns)")
for i in sel[:8]:
    print("  " + " ".join("%7.0f" % (x - t0) for x in b[i, :7]))
d = b[sel]
print("medians [ns]: MMA waits for TMEM %.0f | first MMA -> commit %.0f | epilogue waits for accumulators %.0f | drain TMEM %.0f | "
      "finish stores %.0f | tile period %.0f" % (np.median(d[1:, 1] - d[1:, 0]), np.median(d[:, 2] - d[:, 1]), np.median(d[:, 4] - d[:, 3]),
                                                 np.median(d[:, 5] - d[:, 4]), np.median(d[:, 6] - d[:, 5]), np.median(np.diff(d[:, 0]))))
