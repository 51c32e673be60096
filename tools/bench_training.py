"""Times one training step (config 5): 640 wide-2D conditions, teacher labels.  python tools/bench_training.py [n]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import CrnnTrainer, synthetic_labels

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    gold = os.path.join(ROOT, "tests", "golden")
    a = np.load(os.path.join(gold, "conditions.npz"))["training_wide_2D"][:n]
    sur = Surrogate(ModelSet.from_packed(os.path.join(gold, "containers", "LLNL.npz"), "Eoff"))
    teacher = ModelSet.from_packed(os.path.join(gold, "containers", "LLNL.npz"), "Eoff", "Eoff_wide").crnn
    batch = synthetic_labels(sur, teacher, a[:, 0].astype(np.float32), (a[:, 1] * 1e5).astype(np.float32))
    tr = CrnnTrainer(batch)
    kat = np.load(os.path.join(gold, "converter_kat.npz"))
    p = (torch.tensor(kat["LLNL_Eoff_wide/updated_p"]) + 0.05 * torch.randn(189, generator=torch.Generator().manual_seed(0))).requires_grad_(True)
    losses = []
    for _ in range(3):
        losses.append(tr.step(p)[0])
    torch.cuda.synchronize(); t0 = time.time()
    K = 10
    for _ in range(K):
        losses.append(tr.step(p)[0])
    torch.cuda.synchronize(); dt = (time.time() - t0) / K
    # split: forward vs adjoint
    w = [x.detach().numpy() for x in tr.converter(p.detach())]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); crnn, res = tr.forward(*w); e[1].record(); tr.packed_loss_grad(*w); e[2].record(); torch.cuda.synchronize()
    # host-side share: the step without waiting for the device work of the previous one
    t0 = time.time()
    for _ in range(K):
        pc = p.detach().to("cpu", torch.float32).requires_grad_(True)
        w_in, w_b, w_out = tr.converter(pc)
        (gp,) = torch.autograd.grad((w_in, w_b, w_out), pc, (torch.ones(11, 9), torch.ones(9), torch.ones(9, 9)))
    host_conv = (time.time() - t0) / K
    st = res.stats.double()
    print(json.dumps({"n": n, "step_ms": dt * 1e3, "samples_per_s": n / dt, "forward_ms": e[0].elapsed_time(e[1]),
                      "forward_plus_adjoint_ms": e[1].elapsed_time(e[2]), "converter_fwd_bwd_host_ms": host_conv * 1e3,
                      "forward_attempts_mean": float((st[0] + st[1]).mean()), "forward_attempts_max": float((st[0] + st[1]).max()),
                      "forward_method": tr.forward_method, "losses": losses[:3] + losses[-2:]}))

if __name__ == "__main__":
    main()
