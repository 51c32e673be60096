"""Builds the in-tree CUDA shared library (sm_100a only) with nvcc.  No JIT cache: the .so lives next
to the package so it travels with a snapshot of the repository."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcrnn_pfr_b200.so")
SOURCES = ["capi.cu"]
HEADERS = ["sweep_order.cuh", "adjoint_phases.cuh", "integrate_lanes.cuh", "crnn_device.cuh", "integrate_explicit.cuh", "integrate_taylor.cuh", "fastmath.cuh", "integrate_rodas.cuh", "integrate_rodas_coop.cuh", "adjoint.cuh", "integrate_dopri5.cuh", "mlp.cuh", "mlp_tc.cuh", "mlp_train.cuh", "../../include/crnn_pfr.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=true",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS) or os.path.getmtime(__file__) > t


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """Compile the library.  `defines`/`out` build a tuning variant next to the default one."""
    if out is None and not force and not needs_build():
        return LIB
    out = out or LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out, *[os.path.join(CSRC, s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log" if out == LIB else os.path.basename(out) + ".log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode != 0:
        sys.stderr.write(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed; see build.log")
    return out


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
