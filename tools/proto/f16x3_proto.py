"""CPU prototype of the f16x3 operand split of the predictor MLPs (tools only).

x = hi + lo' * 2^-11 with hi = rn_f16(x), lo' = rn_f16((x - hi) * 2^11); a*b ~= hi*hi + 2^-11 (hi*lo' + lo'*hi).
Compares the representation error of that split against the tf32 split the tensor-core path ships (hi = rn_tf32(x),
lo = rn_tf32(x - hi)) on the shipped weights and on LHS inputs, with products accumulated in float64 (the accumulation
behaviour of the tensor core is the same for both), and reports the activation / weight range against float16's.
"""
import sys
import numpy as np

sys.path.insert(0, ".")
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet


def rn_tf32(x):
    u = np.asarray(x, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def split_tf32(x):
    hi = rn_tf32(x)
    return hi.astype(np.float64), rn_tf32((x - hi).astype(np.float32)).astype(np.float64), 1.0


def split_f16(x):
    hi = x.astype(np.float16)
    lo = ((x - hi.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64), 2.0 ** -11


def layer(split, h, W, b, relu=True):
    ah, al, sa = split(h)
    wh, wl, sw = split(W)
    out = ah @ wh.T + sa * (al @ wh.T) + sw * (ah @ wl.T)
    out = (out + b).astype(np.float32)
    return np.maximum(out, 0) if relu else out


def run(split, mlp, X):
    h = np.maximum((X.astype(np.float32) @ mlp.w[0].T + mlp.b[0]).astype(np.float32), 0)
    stats = [float(h.max())]
    for i in (1, 2):
        h = layer(split, h, mlp.w[i], mlp.b[i])
        stats.append(float(h.max()))
    return layer(split, h, mlp.w[3], mlp.b[3], relu=False), stats


def exact(mlp, X):
    h = np.maximum(X.astype(np.float64) @ mlp.w[0].astype(np.float64).T + mlp.b[0], 0)
    for i in (1, 2):
        h = np.maximum(h @ mlp.w[i].astype(np.float64).T + mlp.b[i], 0)
    return h @ mlp.w[3].astype(np.float64).T + mlp.b[3]


rng = np.random.default_rng(0)
for mech in ("LLNL", "JetSurf", "NUIG"):
    ms = ModelSet.from_packed(f"tests/golden/containers/{mech}.npz", "Eon")
    for name, mlp in (("time", ms.time_mlp), ("temp", ms.temp_mlp)):
        X = rng.random((2048, mlp.in_dim))
        ref = exact(mlp, X)
        wmax = max(float(np.abs(w).max()) for w in mlp.w[1:])
        wmin = min(float(np.abs(w[w != 0]).min()) for w in mlp.w[1:])
        for sname, split in (("tf32x3", split_tf32), ("f16x3", split_f16)):
            out, stats = run(split, mlp, X)
            err = np.abs(out - ref)
            print(f"{mech} {name} {sname}: max abs err {err.max():.3e} mean {err.mean():.3e} | act max {stats} | |w| in [{wmin:.2e}, {wmax:.2e}]")
