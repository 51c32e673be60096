#!/usr/bin/env python
"""Contract benchmark: PFR trajectories per second on the synthetic 1M-condition 4-D Latin-hypercube sweep.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the reference's CPU path, oracle port, all host cores)

One "step" = one pass of the surrogate hot path over one batch of 2^20 conditions per GPU (weak scaling):
inlet concentration -> temperature MLP + time MLPs -> enforce_strict / idx_cut -> adaptive Rosenbrock
integration -> outlet species [9, n] (+ the final gather when N > 1).  Headline workload: LLNL Eon (the
coupled CRNN + temperature-profile MLP path), float64 state.  Integrator of the headline: the explicit knot-limited
fast path (BS23, one step per knot interval of the MLP grid; conditions it flags as stiff fall back to the Rosenbrock
kernel) at a tolerance tight enough that its outlet error is below that of the 6-stage Rosenbrock method RODAS4 at the
reference's 1e-6 -- all errors are measured in the run and reported under "accuracy"; the Rosenbrock methods (ROS3 at
1e-7, RODAS4 at 1e-6) are timed under "variants".  The LLNL Eoff (isothermal, free-stepping) sweep is timed as well and
reported under "variants": with its explicit fast path (DP54 at 1e-7) and with RODAS4 at 1e-6.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "containers")
METRIC = "PFR trajectories/sec (1M LHS conditions)"

# FP64-pipe instruction counts per unit of ALGORITHMIC work (DESIGN.md "Roofline accounting"), i.e. what one thread
# that owned a whole condition would have to execute with the kernel's own log/exp (fastmath.cuh); 1 instruction
# = 2 flop, so achieved / measured-DFMA-peak is the FP64-pipe utilisation a redundancy-free schedule would show.
# The 3-lane kernel executes ~35 % more than this (replicated control flow, Gauss-Jordan instead of LU).
FP64_LOG, FP64_EXP, FP64_RCP = 8, 9, 5                        # table-driven log / exp (256-entry tables, degree 4), MUFU.RCP64H + 2 Newton steps
FP64_RHS = 2 * 81 + 9 * FP64_LOG + 9 * FP64_EXP + 27           # two 9x9 mat-vecs, 9 log + 9 exp, clamp compares
FP64_RHS_T = FP64_LOG + FP64_RCP + 27                          # on a T ramp: ln T, 1/T, kT_j = lnA - Ea/RT + b lnT
# the explicit kernels (bs23 / dp54): exponents pre-scaled (exp = 8 instructions), only the state clamp compares on the FP64 pipe
# (18 DSETP; the exponent / output clamps are decided on the integer pipe), kT_j as two FMAs per reaction
FP64_RHS_EXPL = 2 * 81 + 9 * FP64_LOG + 9 * 8 + 18
FP64_RHS_T_EXPL = FP64_LOG + FP64_RCP + 19
FP64_STEP = (81 + 729 + 81) + (204 + 36 + 9 * FP64_RCP) + 6 * 81 + (90 + 18 + 135 + 15) + 50   # J, LU, 6 solves, stage sums, norm
FP64_STEP_ROS3 = (81 + 729 + 81) + (204 + 36 + 9 * FP64_RCP) + 3 * 81 + 108 + 50                # J, LU, 3 solves, stage sums, norm
FP64_STEP_BS23 = 9 + 45 + 27 + 63                                                              # stage arguments / solution / error combinations of the three stages, norm
FP64_STEP_DP54 = 9 * (21 + 6 + 7 + 4)                                                          # stage arguments, error combination, norm
# Taylor-4: one coefficient evaluation = one log / exp set, eight mat-vecs, the kT series terms, 1/Y, the log / exp series recurrences
FP64_TAYLOR4_EVAL = 8 * 81 + 9 * FP64_LOG + 9 * FP64_EXP + 27 + 54 + 27 + 198
FP64_STEP_TAYLOR4 = 36 + 63                                                                      # Horner, error norm
STEP_INSTR = {"rodas4": FP64_STEP, "ros3": FP64_STEP_ROS3, "bs23": FP64_STEP_BS23, "dp54": FP64_STEP_DP54, "taylor4": FP64_STEP_TAYLOR4}


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled every 100 ms while the timed region runs.  In-process NVML
    (nvidia_ml_py) on a thread, initialised before the first step: an `nvidia-smi -lms` child process was measured to stall
    kernel launches for ~130 ms once, early in its life, which showed up as one slow step in the timed region."""
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.h, self.stop = index, [], None, threading.Event()

    def __enter__(self):
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._sample()
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        except Exception:
            self.h = None
        return self

    def _sample(self):
        nv = self.nv
        sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        self.rows.append((sm, reasons))

    def _run(self):
        while not self.stop.wait(0.1):
            try:
                self._sample()
            except Exception:
                break

    def __exit__(self, *exc):
        self.stop.set()
        if self.h is not None:
            self.th.join(timeout=2)
        return False

    def summary(self):
        sm = [r[0] for r in self.rows[1:]] or [r[0] for r in self.rows]
        reasons = set()
        for _, bits in self.rows[1:]:
            for b, nme in self.BITS.items():
                if bits & b:
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": getattr(self, "mx", None), "reasons": sorted(reasons),
                "samples": len(sm), "source": "NVML (nvidia_ml_py), in-process, every 100 ms during the timed steps"}


def _flops(stats, energy_on, method="rodas4"):
    """Algorithmic FP64 work of one integrator launch from its per-trajectory counters [3, n]."""
    import torch
    acc, rej, rhs = (stats[i].to(torch.float64).sum().item() for i in range(3))
    explicit = method in ("bs23", "dp54")
    per_rhs = FP64_TAYLOR4_EVAL if method == "taylor4" else (FP64_RHS_EXPL if explicit else FP64_RHS)
    per_rhs_t = FP64_RHS_T_EXPL if explicit else FP64_RHS_T
    instr = rhs * (per_rhs + (per_rhs_t if energy_on else 0)) + (acc + rej) * STEP_INSTR[method]
    return 2.0 * instr, {"accepted_mean": acc / stats.shape[1], "rejected_mean": rej / stats.shape[1], "rhs_mean": rhs / stats.shape[1]}


def run_ours(args, emit=print):
    import torch
    import torch.distributed as dist

    from n_hexane_pyrolysis_surrogate_reactor_model_b200 import _lib
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import FAST_TOLERANCE, Surrogate, measure_peaks
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import gather_outlets, lhs_conditions, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 under torchrun (python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_total = args.conditions_per_gpu * world
    Th, Ph, Lh, Uh = lhs_conditions(n_total, seed=13895)
    lo, hi = shard_bounds(n_total, world, rank)
    host = [torch.from_numpy(np.ascontiguousarray(a[lo:hi])).pin_memory() for a in (Th, Ph, Lh, Uh)]
    n = hi - lo
    host_np = [h.numpy() for h in host]
    T, P, L, U = (h.to(dev) for h in host)
    # strong scaling: the SAME 2^20 conditions whatever the number of GPUs (the first conditions_per_gpu rows of the hypercube)
    slo, shi = shard_bounds(args.conditions_per_gpu, world, rank)
    Ts, Ps, Ls, Us = (torch.from_numpy(np.ascontiguousarray(a[slo:shi])).to(dev) for a in (Th, Ph, Lh, Uh))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_steps(fn, steps, warmup):
        out = None
        for _ in range(warmup):
            out = fn()   # same lifetime pattern as the timed loop: the previous result is still alive while the next one is
                         # allocated, so the caching allocator reaches its steady state (two result sets) during warm-up
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.lib().pfr_launch_count()
        marks = [e0]
        e0.record()
        for _ in range(steps):
            t_host = time.perf_counter()
            out = fn()
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
            if os.environ.get("PFR_BENCH_TRACE"):
                print(f"[trace] {getattr(fn, '__name__', 'fn')} host-side {1e3 * (time.perf_counter() - t_host):.1f} ms", file=sys.stderr, flush=True)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        time_steps.each = [marks[j].elapsed_time(marks[j + 1]) for j in range(steps)]   # this rank's individual steps
        return float(ms.item()), out, _lib.lib().pfr_launch_count() - l0

    def error_triple(y, yref):
        e = ((y.double() - yref).abs() / torch.clamp(yref.abs(), min=1e-3)).amax(0)
        return {"max": float(e.max()), "p99": float(torch.quantile(e, 0.99)), "median": float(e.median()), "over_1e-6": int((e > 1e-6).sum())}

    peaks = measure_peaks() if rank == 0 else None
    result, variants = None, {}
    accuracy = None
    bs_r, bs_a = FAST_TOLERANCE["bs23"] if args.bs23_rtol is None else (args.bs23_rtol, args.bs23_atol)
    dp_r, dp_a = FAST_TOLERANCE["dp54"] if args.dp54_tol is None else (args.dp54_tol, args.dp54_tol)
    # (name, variant, MLP arithmetic, integrator, state precision, rtol, atol)
    runs = (("LLNL_Eon", "Eon", "f16x3", "bs23", args.precision, bs_r, bs_a),
            ("LLNL_Eoff", "Eoff", "f16x3", "dp54", args.precision, dp_r, dp_a),
            ("LLNL_Eon_bs23_loose", "Eon", "f16x3", "bs23", 64, 1e-6, 1e-12),
            ("LLNL_Eon_bs23_round1_setting", "Eon", "f16x3", "bs23", 64, 1e-8, 1e-8),
            ("LLNL_Eon_fast32", "Eon", "f16x3", "bs23", 32, 1e-7, 1e-7),
            ("LLNL_Eoff_loose", "Eoff", "f16x3", "dp54", 64, 1e-7, 1e-7),
            ("LLNL_Eoff_fast32", "Eoff", "f16x3", "dp54", 32, 1e-7, 1e-7),
            ("LLNL_Eoff_rodas4", "Eoff", "f16x3", "rodas4", 64, args.rtol, args.atol),
            ("LLNL_Eon_ros3", "Eon", "f16x3", "ros3", 64, args.ros3_tol, args.ros3_tol),
            ("LLNL_Eon_rodas4", "Eon", "f16x3", "rodas4", 64, args.rtol, args.atol),
            ("LLNL_Eon_taylor4", "Eon", "f16x3", "taylor4", 64, bs_r, bs_a),
            ("LLNL_Eon_mlp_tf32x3", "Eon", "tf32x3", "bs23", 64, bs_r, bs_a),
            ("LLNL_Eon_mlp_fp32", "Eon", "fp32", "bs23", 64, bs_r, bs_a))
    if args.headline_only:   # (multi-GPU scaling checks: the headline Eon sweep, the Eoff sweep and the training step only)
        runs = runs[:2]
    sur, sur_key, grids = None, None, None
    for name, variant, mlp_mode, method, prec, rtol, atol in runs:
        if sur_key != (variant, mlp_mode):
            del sur, grids
            torch.cuda.empty_cache()
            sur = Surrogate(ModelSet.from_packed(os.path.join(GOLD, "LLNL.npz"), variant), device=dev, mlp_mode=mlp_mode)
            sur_key, grids = (variant, mlp_mode), None
        headline = name == "LLNL_Eon"
        kw = dict(method=method, precision=prec, rtol=rtol, atol=atol)

        def step_device():
            r = sur.sweep(T, P, L, U, staged=False, **kw)
            r.y_all = gather_outlets(r.y, n_total, as_blocks=True)
            return r

        def step_e2e():
            # the public host-buffer call: numpy conditions in, numpy outlets + status of this rank's shard out
            # (page-locked staging inside Surrogate.sweep_host), plus the device-side gather of the whole job
            y_host, st_host, r = sur.sweep_host(*host_np, staged=False, **kw)
            r.y_all = gather_outlets(r.y, n_total, as_blocks=True)
            r.y_host, r.status_host = y_host, st_host
            return r

        def step_strong():
            r = sur.sweep(Ts, Ps, Ls, Us, staged=False, **kw)
            r.y_all = gather_outlets(r.y, args.conditions_per_gpu, as_blocks=True)
            return r

        full = headline or name == "LLNL_Eoff"
        steps = args.steps if headline else max(1, min(args.steps, 3))
        with ClockSampler(local) as clk:
            ms, res, launches = time_steps(step_device, steps, args.warmup if headline else 3)
        each = list(time_steps.each)
        kms = sur.integrator_ms()          # mean over the timed steps: event pairs the sweep handle records around the integrator launch
        bad = int((res.status != 0).sum().item())
        flops, work = _flops(res.stats, variant == "Eon", method)
        entry = {"value": n_total * steps / (ms * 1e-3), "ms_per_step": ms / steps, "failed_trajectories": bad, "work_per_trajectory": work,
                 "step_ms_each": each, "integrator_ms": kms, "integrator_share_of_step": kms / (ms / steps),
                 "stiff_fallbacks": res.stiff_fallbacks, "mlp_arithmetic": mlp_mode, "integrator": method, "state_dtype": f"f{prec}",
                 "rtol": rtol, "atol": atol}
        if prec == 64:
            entry["integrator_fp64_tflops"] = flops / (kms * 1e-3) / 1e12
        if full:
            ms_e2e, res2, _ = time_steps(step_e2e, steps, args.warmup)
            entry["e2e"], entry["e2e_step_ms_each"] = n_total * steps / (ms_e2e * 1e-3), list(time_steps.each)
            ms_s, _, _ = time_steps(step_strong, steps, 3)
            entry["strong"] = {"conditions_total": args.conditions_per_gpu, "value": args.conditions_per_gpu * steps / (ms_s * 1e-3),
                               "ms_per_step": ms_s / steps, "conditions_per_gpu": shi - slo}
        # outlet deviation from the tight-tolerance solution of the Rosenbrock kernel (its parity with the converged CPU solution is what
        # tests/test_gpu_parity.py establishes) on every 16th condition of this rank's shard, on the same grids
        if rank == 0 and mlp_mode == "f16x3" and method in ("bs23", "taylor4", "dp54", "ros3", "rodas4"):
            if grids is None:
                sel = torch.arange(0, n, 16, device=dev)
                c0 = sur.inlet_concentration(T[sel], P[sel])
                if variant == "Eon":
                    tf, _ = sur.time_grid(T[sel], P[sel])
                    Tp = sur.temp_profile(T[sel], P[sel])
                    _, te = sur.time_grid(T[sel], P[sel], L[sel], U[sel], want_grid=False, want_end=True)
                    g = dict(tgrid=tf, Tprof=Tp, idx_end=sur.idx_cut(tf, te))
                else:
                    _, te = sur.time_grid(T[sel], P[sel], L[sel], U[sel], want_grid=False, want_end=True)
                    g = dict(t_end=te)
                yref = sur.integrate(T[sel], c0, method="rodas4", rtol=1e-11, atol=1e-11, **g).y.clone()
                grids = (sel, c0, g, yref)
            sel, c0, g, yref = grids
            y = sur.integrate(T[sel], c0, stiff_fallback=None, **g, **kw).y
            entry["accuracy"] = dict(error_triple(y, yref), conditions=int(sel.numel()))
            # the sweep's own outlets at the same conditions: the one-call pipeline must reproduce the staged kernels bit for bit
            entry["accuracy"]["one_call_equals_staged"] = bool(torch.equal(res.y[:, sel], y))
        variants[name] = entry
        if headline:
            result = dict(entry=entry, clk=clk.summary(), launches=launches, flops=flops, kms=kms, steps=steps, ms=ms,
                          mean_outlet_knot=float(res.idx_cut.double().mean()))
            if "accuracy" in entry:   # rank 0
                accuracy = {"reference_solution": "RODAS4 at rtol = atol = 1e-11 on the same grids", "conditions": entry["accuracy"]["conditions"],
                            "error": "max over species of |y - y_ref| / max(|y_ref|, 1e-3 mol/m3) at the outlet",
                            "parity_bound": 1e-6, "headline": entry["accuracy"]}
    for nm in ("LLNL_Eon_bs23_loose", "LLNL_Eon_bs23_round1_setting", "LLNL_Eon_fast32", "LLNL_Eon_taylor4", "LLNL_Eon_ros3", "LLNL_Eon_rodas4",
               "LLNL_Eoff", "LLNL_Eoff_loose", "LLNL_Eoff_fast32", "LLNL_Eoff_rodas4"):
        if accuracy is not None and "accuracy" in variants.get(nm, {}):
            accuracy[nm] = dict(variants[nm]["accuracy"], rtol=variants[nm]["rtol"], atol=variants[nm]["atol"])
    del sur, grids
    torch.cuda.empty_cache()

    # the other mechanisms of config 3 and the reference-behaviour integrator, device-resident timing only
    for mech, variant, method, prec in () if args.headline_only else (
            ("JetSurf", "Eon", "bs23", 64), ("JetSurf", "Eoff", "dp54", 64), ("NUIG", "Eon", "bs23", 64),
            ("NUIG", "Eoff", "dp54", 64), ("LLNL", "Eoff", "dopri5", 32), ("LLNL", "Eon", "rodas4", 32)):
        sur = Surrogate(ModelSet.from_packed(os.path.join(GOLD, f"{mech}.npz"), variant), device=dev)
        rt, at = {"bs23": (bs_r, bs_a), "dp54": (dp_r, dp_a)}.get(method, (args.rtol, args.atol))
        kw2 = dict(method=method, precision=prec, rtol=rt, atol=at)
        ms, res, _ = time_steps(lambda: gather_outlets(sur.sweep(T, P, L, U, **kw2).y, n_total, as_blocks=True), 2, 3)   # 3 warm-ups: allocator steady state
        r = sur.sweep(T, P, L, U, **kw2)
        name = f"{mech}_{variant}" + ("" if prec == 64 else f"_{method}_f{prec}")
        variants[name] = {"value": n_total * 2 / (ms * 1e-3), "ms_per_step": ms / 2, "failed_trajectories": int((r.status != 0).sum().item()),
                          "integrator": method, "state_dtype": f"f{prec}", "rtol": rt, "atol": at}
        del sur, r, res
        torch.cuda.empty_cache()
    variants["train_step_WIDE_Eoff"] = training_variant(dev, world, rank, time_steps)
    variants["mlp_train_step_2D"] = mlp_training_variant(dev, time_steps)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    e = result["entry"]
    peak = peaks["dfma_flops"]
    # DRAM traffic of the dominant kernel.  Derived in this run from the outlet knots the kernel walked to: per condition it reads
    # T0, c0, idx_end (12 B) and two float32 grid values for every knot up to idx_cut + 2 (the two-knot look-ahead), and writes 9
    # float64 outlets, status and three counters (88 B).  Measured with ncu on the same kernel (profiles/r02b_ncu_full_bs23_dp54_fp64.txt,
    # 131 072 conditions): 4.02 KB read + 0.19 KB written per condition, L2 sector hit rate 71 % (round 1, gathered reads: 74.8 KB).
    traffic_model = n * (12 + 88 + 8.0 * (result["mean_outlet_knot"] + 3))
    line = {
        "metric": METRIC, "value": e["value"], "unit": "trajectories/s", "n_gpus": world, "steps": result["steps"],
        "warmup": args.warmup, "ms_per_step": e["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if args.precision == 64 else "f32", "data": "synthetic",
        "config": {"workload": "LLNL Eon CRNN + LLNL_2D temperature MLP + LLNL_4D_time_on MLP; 4-D Latin hypercube "
                               "(T 870-1150 K, P 1-3 bar, L 0.5-1 m, u0 2.5-5 m/s), scipy qmc seed 13895",
                   "conditions_per_gpu": args.conditions_per_gpu, "conditions_total": n_total,
                   "integrator": "bs23: explicit Bogacki-Shampine 3(2), adaptive, knot-limited steps, one thread per condition (PFR_METHOD_BS23), "
                                 "CRNN coefficients through uniform registers (LDCU + DFMA R,R,UR,R), "
                                 "Rosenbrock (ROS3) fallback for conditions flagged stiff (device-side list); run at the PARITY-CERTIFIED "
                                 "setting: the loosest (rtol, atol) at which every sampled condition is within 1e-6 of the tight-tolerance "
                                 "solution (see accuracy.headline; looser / float32 settings are timed under variants)",
                   "api": "Surrogate.sweep -> pfr_sweep_run: the whole hot path as one C-ABI call (ordering, inlet, 3 MLP passes, idx_cut, "
                          "integrator, fallback), nothing waits for the host",
                   "mlp_arithmetic": "tcgen05 tensor cores, error-compensated three-product split with float16 operand pairs (hi + 2^-11 lo'), "
                                     "four float32 TMEM accumulators per tile (float32-accurate: 1.2e-6 vs the FP32-FFMA path, the same as the "
                                     "3xTF32 split it replaces at half the operand bytes; PFR_MLP_F16X3)",
                   "rtol": bs_r, "atol": bs_a, "weights": "trained reference containers (tests/golden/containers)",
                   "l2": "per-step working set (6.4 KB of grids per condition, 6.7 GB per GPU) exceeds the 126 MB L2",
                   "parallelism": f"conditions sharded over {world} rank(s); final all_gather_into_tensor of [9,n] outlets only"},
        "e2e": {"value": e["e2e"], "unit": "trajectories/s", "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": (72 if args.precision == 64 else 36) * n + 4 * n,
                "api": "Surrogate.sweep_host (numpy in, numpy out; per rank: its shard of conditions in, its outlets + status out)"},
        "strong": e["strong"],
        "gpu_launches": int(result["launches"]),
        "clocks": result["clk"],
        "roofline": {"bound": "fp64_pipe", "kernel": "bs23_kernel<double,ramp>", "achieved": result["flops"] / (result["kms"] * 1e-3) / 1e12,
                     "peak": peak / 1e12, "unit": "TFLOP/s", "frac": result["flops"] / (result["kms"] * 1e-3) / peak,
                     "peak_source": "pfr_measure_peaks(): dependent-free DFMA loop measured in this run (MEASURED_PEAKS.json holds no FP64 figure)",
                     "kernel_ms": result["kms"], "kernel_ms_source": "CUDA event pairs the sweep handle records around the kernel launch inside the timed steps (mean over the steps)",
                     "kernel_share_of_step": e["integrator_share_of_step"],
                     "traffic": traffic_model, "traffic_per_condition": traffic_model / n,
                     "traffic_source": "derived in this run from the outlet knots walked (inputs 12 B, 8 B per knot up to idx_cut + 2, results 88 B); "
                                       "ncu on the same kernel: 4.2 KB per condition at 2^17 conditions (L2 hit rate 71 %, "
                                       "profiles/r02b_ncu_full_bs23_dp54_fp64.txt), 10.8 KB at 2^19 (57 %, profiles/r02p_ncu_full_bs23_fp64.txt); "
                                       "round 1, reads gathered through a permutation: 74.8 KB",
                     "flop_model": f"2 flop per FP64-pipe instruction of the algorithm as shipped: RHS {FP64_RHS_EXPL} (+{FP64_RHS_T_EXPL} on a T ramp; "
                                   f"two 9x9 mat-vecs, 9 log + 9 exp at 8 instructions each, 18 clamp compares), BS23 step overhead {FP64_STEP_BS23} "
                                   f"(stage sums, error norm); the Rosenbrock variants: RHS {FP64_RHS} (+{FP64_RHS_T}), step ROS3 {FP64_STEP_ROS3}, "
                                   f"RODAS4 {FP64_STEP}; see DESIGN.md. (Round 1 / the first half of round 2 counted 421 (+43) per RHS: the log / exp "
                                   f"were 11 + 10 instructions then, so fractions across rounds compare pipe utilisation, not speed.)"},
        "accuracy": accuracy,
        "peaks_measured": {"ffma_tflops": peaks["ffma_flops"] / 1e12, "dfma_tflops": peaks["dfma_flops"] / 1e12, "mufu_tops": peaks["mufu_ops"] / 1e12},
        "variants": variants,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def training_variant(dev, world, rank, time_steps):
    """Config 5: one CRNN training step (batched forward + adjoint gradient + 189-float all-reduce + clip + AdamW) over the 640
    training conditions of sampling_case_wide_2D.csv (train_test_split seed 42 is immaterial to the timing: the first 640 rows),
    sharded over the ranks; teacher-generated labels (the reference's Cantera labels are not shipped)."""
    import torch
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import shard_bounds
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.training import CrnnTrainer, synthetic_labels

    gold = os.path.join(ROOT, "tests", "golden")
    a = np.load(os.path.join(gold, "conditions.npz"))["training_wide_2D"][:640]
    lo, hi = shard_bounds(len(a), world, rank)
    sur = Surrogate(ModelSet.from_packed(os.path.join(GOLD, "LLNL.npz"), "Eoff"), device=dev)
    teacher = ModelSet.from_packed(os.path.join(GOLD, "LLNL.npz"), "Eoff", "Eoff_wide").crnn
    batch = synthetic_labels(sur, teacher, a[lo:hi, 0].astype(np.float32), (a[lo:hi, 1] * 1e5).astype(np.float32))
    tr = CrnnTrainer(batch)
    kat = np.load(os.path.join(gold, "converter_kat.npz"))
    p = (torch.tensor(kat["LLNL_Eoff_wide/updated_p"]) + 0.05 * torch.randn(189, generator=torch.Generator().manual_seed(0))).requires_grad_(True)
    losses = []
    ms, _, launches = time_steps(lambda: losses.append(tr.step(p)[0]), 5, 3)
    return {"value": len(a) * 5 / (ms * 1e-3), "unit": "training samples/s", "ms_per_step": ms / 5, "conditions": len(a),
            "kernel_launches_per_step": launches / 5, "loss_first": losses[0], "loss_last": losses[-1],
            "reference": "WIDE_Eoff_surrogate_model_training.py: ~0.225 s per sample per CPU core (SURVEY 6), batch size 1"}


def _mlp_training_problem():
    """Temperature-MLP training data of temp_profile_model_training_2D.py's shape: the 800 (T, P) rows of sampling_case_2D.csv with
    labels from the shipped LLNL temperature MLP (teacher; the Cantera label files are absent), evaluated by the device
    trainer's own forward kernels."""
    import torch
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.mlp_training import INPUT_SCALE, MlpDataset, MlpTrainer, TEMP_2D_SETTINGS
    a = np.load(os.path.join(ROOT, "tests", "golden", "conditions.npz"))["training_2D"]
    teacher = ModelSet.from_packed(os.path.join(GOLD, "LLNL.npz"), "Eon").temp_mlp
    sc = INPUT_SCALE[2]
    x = ((a - sc[0]) / (sc[1] - sc[0])).astype(np.float32)
    y = MlpTrainer(teacher.w, teacher.b, TEMP_2D_SETTINGS).forward(torch.from_numpy(x)).double().cpu().numpy()
    return MlpDataset(a, y * (teacher.out_max - teacher.out_min) + teacher.out_min)


def mlp_training_variant(dev, time_steps):
    """optimisation steps of the temperature-predictor MLP (batch 32, Adam) on the device."""
    import torch
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.mlp_training import MlpTrainer, TEMP_2D_SETTINGS, initial_parameters
    ds = _mlp_training_problem()
    xt, yt = (torch.from_numpy(v).to(dev) for v in ds.parts["training"])
    w0, b0 = initial_parameters(2, seed=0)
    tr = MlpTrainer(w0, b0, TEMP_2D_SETTINGS, dev)
    loss = torch.zeros((), device=dev)
    batches = [(xt[i:i + 32].contiguous(), yt[i:i + 32].contiguous()) for i in range(0, 640, 32)]
    first = []

    def epoch():
        for bx, by in batches:
            tr.step(bx, by, 1e-3, loss)
        first.append(float(loss)) if len(first) < 1 else None
        return loss
    ms, _, launches = time_steps(epoch, 5, 3)
    return {"value": 5 * len(batches) / (ms * 1e-3), "unit": "optimisation steps/s (batch 32)", "ms_per_step": ms / (5 * len(batches)),
            "kernel_launches_per_step": launches / (5 * len(batches)), "loss_after_first_epoch": first[0], "loss_last": float(loss),
            "reference": "temp_profile_model_training_2D.py:145-160 (torch, batch 32, Adam 1e-3); torch CPU timing under cpu_baseline"}


def cpu_baseline(args, variant="Eon", sample=None):
    """The oracle's restatement of the reference drivers (torch CPU ops, per-condition Python loop) on all host cores."""
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
    from oracle import c_oracle as CO
    from oracle import reference_driver as D
    from oracle import reference_path as R

    cores = os.cpu_count() or 1
    sample = sample or 96 * cores      # ~10-30 s of CPU work at ~10 trajectories/s/core
    T, P, L, U = lhs_conditions(args.conditions_per_gpu, seed=13895)
    idx = np.arange(0, len(T), len(T) // sample)[:sample]
    ms = ModelSet.from_packed(os.path.join(GOLD, "LLNL.npz"), variant)
    t0 = time.time()
    y = D.sweep_parallel(ms, T[idx], P[idx], L[idx], U[idx], cores)
    dt = time.time() - t0
    failed = int(np.isnan(y).any(axis=1).sum())
    # trajectories the port COMPLETED per second: a condition on which torchdiffeq's asserts would fire (`underflow in dt`; the
    # reference script aborts there) is attempted, counted under "failed" and not counted as work done
    out = {"value": (sample - failed) / dt, "unit": "trajectories/s", "cores": cores, "kind": "port", "attempted": sample,
           "sample": f"{sample} of the {len(T)} LHS conditions (every {len(T) // sample}-th), LLNL {variant}, oracle/reference_driver.py "
                     f"(torch CPU float32 ops, dopri5 1e-6/1e-6, 801 outputs, per-condition loop) over {cores} processes",
           "seconds": dt, "failed": failed}
    # for context: the same algorithm as compiled C (oracle/crnn_oracle.c), ODE part only, all cores
    m = min(len(T), 8192)
    sel = np.arange(0, len(T), len(T) // m)[:m]
    tm = R.MLPParams(ms.time_mlp.w, ms.time_mlp.b, ms.time_mlp.out_min, ms.time_mlp.out_max)
    pm = R.MLPParams(ms.temp_mlp.w, ms.temp_mlp.b, ms.temp_mlp.out_min, ms.temp_mlp.out_max)
    t0 = time.time()
    tg = R.time_grid(tm, T[sel], P[sel], np.full(m, 1.0, np.float32), np.full(m, 2.5, np.float32))
    ts = R.time_grid(tm, T[sel], P[sel], L[sel], U[sel])
    Tp = R.temp_profile(pm, T[sel], P[sel])
    t_mlp = time.time() - t0
    rep = np.array([R.eon_idx_cut(tg[i], ts[i, -1]) for i in range(m)], np.int32)
    t0 = time.time()
    CO.dopri5_batch(tg, Tp, R.inlet_concentration(T[sel], P[sel]), ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out, report=rep, nthreads=cores)
    t_ode = time.time() - t0
    try:   # the reference's own torch operators, one epoch of the temperature-MLP training loop on the CPU
        import torch
        from n_hexane_pyrolysis_surrogate_reactor_model_b200.mlp_training import initial_parameters
        from oracle import mlp_training_reference as O
        rng = np.random.default_rng(0)
        w0, b0 = initial_parameters(2, seed=0)
        bt = [(rng.random((32, 2), dtype=np.float32), rng.random((32, 800), dtype=np.float32)) for _ in range(20)]
        O.run_steps(w0, b0, bt[:2], [1e-3] * 2)
        t0 = time.time()
        O.run_steps(w0, b0, bt, [1e-3] * 20)
        out["mlp_training_torch_cpu"] = {"steps_per_s": 20 / (time.time() - t0), "threads": torch.get_num_threads(), "batch": 32,
                                         "what": "oracle/mlp_training_reference.py: nn.Linear / MSELoss / Adam of temp_profile_model_training_2D.py"}
    except Exception as exc:   # noqa: BLE001
        out["mlp_training_torch_cpu"] = {"unavailable": repr(exc)}
    out["c_port_for_context"] = {"ode_trajectories_per_s": m / t_ode, "batched_torch_mlp_trajectories_per_s": m / t_mlp, "sample": m,
                                 "note": "compiled C dopri5 (float32, 801 outputs) on all cores + batched torch-CPU MLPs; not what the reference runs"}
    return out


def run_reference(args, emit=print):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals, cb = [], None
    for _ in range(args.warmup and 1):
        cpu_baseline(args, sample=4 * cores)
    for _ in range(max(1, args.steps)):
        cb = cpu_baseline(args, sample=48 * cores)
        vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "trajectories/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * 48 * cores / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "LLNL Eon; the same 4-D Latin hypercube; bounded sample per step (see cpu_baseline.sample)",
                       "note": "the reference scripts cannot run (torchdiffeq, cantera and the label files are absent): oracle port timed"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--conditions-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--precision", type=int, default=64, choices=[32, 64])
    ap.add_argument("--rtol", type=float, default=1e-6)
    ap.add_argument("--atol", type=float, default=1e-6)
    ap.add_argument("--ros3-tol", type=float, default=1e-7, help="rtol = atol of the 3-stage Rosenbrock method on the Eon path")
    ap.add_argument("--bs23-rtol", type=float, default=None, help="rtol of the explicit fast path on the Eon path (default: FAST_TOLERANCE)")
    ap.add_argument("--bs23-atol", type=float, default=1e-12, help="atol that goes with --bs23-rtol")
    ap.add_argument("--dp54-tol", type=float, default=None, help="rtol = atol of the explicit fast path on the isothermal (Eoff) path (default: FAST_TOLERANCE, rtol 1e-7 / atol 1e-10)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="time the headline sweeps and the training step only (no secondary variants)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: whatever libraries print while the run is going on (NCCL's version banner, ...)
    # is sent to stderr by pointing file descriptor 1 there until the line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    try:
        (run_reference if args.impl == "reference" else run_ours)(args, lines.append)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for ln in lines:
        print(ln, flush=True)


if __name__ == "__main__":
    main()
