#!/usr/bin/env python
"""bench.py -- particle-updates/sec of the fused PLS Langevin step (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c4|c3|c2|c5] [--grid RxC] [--gram generated|staged|cached|auto]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one Langevin step over all J particles of the workload:
  C4 (default, the configuration the metric is quoted on): N=1,000,000 D=8 ARD, M=1024, J=4096, Gaussian cost.
At --gpus N the NAMED shape is strong-scaled: the J = 4096 particles are split over the N GPUs (1 x N grid, J_local = J / N,
replicated X, y, Z, V~, lambda; no per-step communication; Philox noise keyed on the global particle index), `scaling` is
"strong" and `config.J_global` stays 4096.  `--grid RxC` shards the training rows R ways as well (NCCL all-reduce of the
M x J_local gradient per step over the row group).

The JSON line carries: value (device-timed, inputs resident in HBM), e2e (through the reference-facing API with pinned
HOST buffers, copies inside the timed region), roofline (FP64 tensor, live CUDA-event timing of every contraction launch:
pls_profile_begin / pls_profile_end), cpu_baseline (the reference's own step on the box's host cores, N = 1 only), clocks,
gpu_launches.  Every timed number above is the DEFAULT path (Gram tiles regenerated inside the kernels, nothing N x M in memory).
Informational objects, measured after and outside the headline's timed region:
  N = 1: library_bar (the reference's algebra on cuBLAS on this GPU), gram_cached / gram_staged (the two opt-in Gram modes),
         gaussian_normal_equations (opt-in M x M re-association), c2 / c3 (BASELINE configs 2 and 3: Bernoulli and Poisson costs);
  N > 1: weak (every GPU advancing its own J = 4096 particles: last round's default), row_sharded (the same shape on a
         2 x N/2 grid: rows sharded, NCCL gradient all-reduce timed with CUDA events), row_sharded_parity (a C5-shaped slice,
         M = 4096, D = 16: sharded vs single-GPU particles, max relative error) and, at N = 8, c5 (BASELINE config 5 on a 2 x 4 grid).

`--impl reference` times the reference's own CPU implementation of the path on all host threads: the UNMODIFIED reference
(installed under oracle/_ref by oracle/install_reference.py, run behind oracle/gpytorch_stub because gpytorch is absent) when
it is there, else the oracle's port of it; full N and M, J_c = 256 particles per timed call (oracle/reference_arm.py).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if "reference" in sys.argv and any(a.startswith("--impl") for a in sys.argv):
    # the CPU arm: the reference moves its Gram to the GPU whenever torch.cuda.is_available() (orthonormal.py:43-45,
    # samplers.py:36-40) -- hide the devices before torch initialises CUDA
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import torch  # noqa: E402

METRIC = "PLS particle-updates/sec at N=1M,M=1024,J=4096; % of FP64/HBM roofline"
UNIT = "particle-updates/s"

WORKLOADS = {
    # name: (N, D, M, J, cost)
    "c4": dict(n=1_000_000, d=8, m=1024, j=4096, cost="gaussian", label="C4 UCI-scale synthetic regression N=1M D=8 ARD M=1024 J=4096"),
    "c5": dict(n=20_000_000, d=16, m=4096, j=16384, cost="gaussian",
               label="C5 large synthetic regression N=20M D=16 ARD M=4096 J=16384 (row- and particle-sharded, NCCL gradient all-reduce)"),
    # (eigenvalue_threshold 1e-5: with the default 0 the kept spectrum reaches 1e-17 and the prior term eta / lambda of the update
    # diverges at any usable step size; the reference's experiment configs set a threshold for the same reason)
    "c3": dict(n=100_000, d=1, m=256, j=4096, cost="poisson", threshold=1e-5, label="C3 Poisson f^2 regression N=100k D=1 M=256 J=4096"),
    "c2": dict(n=10_000, d=1, m=64, j=1024, cost="bernoulli", threshold=1e-5, label="C2 1D Bernoulli classification N=10k M=64 J=1024"),
}


def synth(workload: dict, seed: int = 0):
    """Synthetic inputs of SURVEY.md section 8(d), generated on the CPU generator (seed 0) in float64."""
    g = torch.Generator().manual_seed(seed)
    n, d, m = workload["n"], workload["d"], workload["m"]
    if workload["cost"] == "gaussian":
        x = torch.randn(n, d, generator=g, dtype=torch.float64)
        ls = torch.tensor([math.sqrt(d) * (0.75 + 0.5 * k / max(d - 1, 1)) for k in range(d)], dtype=torch.float64)
        y = torch.sin(x.sum(1) / math.sqrt(d)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
        outputscale = 1.0
    else:
        x = torch.linspace(-3, 3, n, dtype=torch.float64).reshape(-1, 1)
        ls = torch.tensor([0.5], dtype=torch.float64)
        curve = 2.0 * torch.sin(1.5 * x.reshape(-1))
        if workload["cost"] == "bernoulli":
            y = torch.bernoulli(torch.sigmoid(curve), generator=g).double()
        else:
            y = torch.poisson(curve**2, generator=g).double()
        outputscale = 1.0
    z_idx = torch.arange(m) if workload["cost"] == "gaussian" else torch.linspace(0, n - 1, m).long()
    return x, y, x[z_idx].clone(), ls, outputscale


def make_pls(workload: dict, x, y, z, ls, outputscale, gradient_reduce=None, gram_cache="auto", **basis_kw):
    import projected_langevin_sampling_b200 as pkg
    from projected_langevin_sampling_b200.projected_langevin_sampling import costs, link_functions as lf

    kernel = pkg.ScaleKernel(pkg.RBFKernel(ard_num_dims=x.shape[1], lengthscale=ls), outputscale=outputscale)
    threshold = basis_kw.pop("eigenvalue_threshold", workload.get("threshold", 0.0))
    basis = pkg.OrthonormalBasis(pkg.PLSKernel(kernel, z), z, x, eigenvalue_threshold=threshold, verbose=False,
                                 gradient_reduce=gradient_reduce, gram_cache=gram_cache, **basis_kw)
    if workload["cost"] == "gaussian":
        cost = costs.GaussianCost(observation_noise=0.01, y_train=y, link_function=lf.IdentityLinkFunction())
    elif workload["cost"] == "bernoulli":
        cost = costs.BernoulliCost(y_train=y, link_function=lf.SigmoidLinkFunction())
    else:
        cost = costs.PoissonCost(y_train=y, link_function=lf.SquareLinkFunction())
    return pkg.PLS(basis, cost)


# ---- clocks ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi samples (200 ms) of SM clock and throttle reasons DURING the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU arm (the reference's own step on the host cores; oracle/reference_arm.py) ---------------------------------------------------
def step_size_of(workload: dict) -> float:
    # SURVEY.md section 8(d) writes eta = 1e-6 for C4; with lambda_min = 3.6e-9 the prior term eta / lambda makes that diverge within a
    # few steps, so the bench uses 1e-9 there (the step's cost does not depend on eta) and says so in `config`
    return 1e-9 if workload["cost"] == "gaussian" else 1e-6


def base_config(workload: dict, world: int, grid: str) -> dict:
    """The part of `config` that both arms print verbatim (the CPU arm runs the same workload on one host)."""
    return {"workload": workload["label"], "N": workload["n"], "D": workload["d"], "M": workload["m"], "J_global": workload["j"],
            "cost": workload["cost"], "step_size": step_size_of(workload), "grid": grid, "n_gpus": world,
            "l2": "per-step working set (the Dc chunk, N x J_local doubles written + read: GBs) exceeds the 126 MB L2; no flush needed"}


def cpu_arm(workload: dict, steps: int, warmup: int) -> dict:
    from oracle import reference_arm

    return reference_arm.run(workload, synth(workload), steps=steps, warmup=warmup, eta=step_size_of(workload))


def cpu_arm_subprocess(workload_name: str, steps: int, warmup: int, timeout_s: int = 600):
    """cpu_baseline of the GPU arm's line: the CPU arm in its OWN process with the GPUs hidden (the reference moves its Gram to
    the GPU whenever torch.cuda.is_available()), on a bounded number of calls."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload_name, "--steps", str(steps),
           "--warmup", str(warmup)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        line = json.loads(r.stdout.strip().splitlines()[-1])
        return line["cpu_baseline"]
    except Exception as exc:  # a reported baseline, never a reason to lose the line
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}


def run_reference_arm(args, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    t0 = time.perf_counter()
    res = cpu_arm(workload, steps=args.steps, warmup=args.warmup)
    cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "measured")}
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step_whole_j_composed"], "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": base_config(workload, world, args.grid or f"1x{world}"),
        "note": ("CPU arm: one host, all threads, no GPU.  Each of the `steps` timed steps is ONE call of the reference's "
                 "PLS.calculate_particle_update at full N and M on J_c = 256 particles (`cpu_baseline.measured.ms_per_call_median`); "
                 "`ms_per_step` / `value` are the whole-J step composed from the measured J-independent and J-linear parts, with the "
                 "J-independent work counted once (see cpu_baseline.sample)"),
        "cpu_baseline": cpu,
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    emit(line)


# ---- library bar: the reference's formulation on the same GPU with library kernels ------------------------------------------
def library_bar_sample(workload: dict, pls, eta: float, steps: int = 3):
    """SURVEY.md section 8(d) "library bar": the reference's own step algebra (orthonormal.py:98-108,151-158) run on this GPU
    with torch fp64 matmul (cuBLAS DGEMM) and the dense Gram of a bounded row block cached on the device -- what moving the
    reference to CUDA unchanged would give.  A 65 536-row block at full M and J is timed with CUDA events and scaled
    linearly to N.  Reported beside the headline; nothing of it runs inside the timed region."""
    from projected_langevin_sampling_b200 import _native as nat, ops

    ctx = nat.context()
    basis = pls.basis
    n_full, j = workload["n"], workload["j"]
    n_s = min(n_full, 65536)
    eng = basis.engine(j)
    k_xz = ops.gram(ctx, eng.kernel_id, eng.xa[:n_s], eng.za, eng.d)  # (N_s, M) dense, cached like the reference's K_zx
    k_zx = k_xz.T.contiguous()
    vt, lam = basis.scaled_eigenvectors, basis.eigenvalues
    y = pls.cost.y_device()[:n_s]
    p = torch.randn(vt.shape[1], j, dtype=torch.float64, device=vt.device)
    xi = torch.randn_like(p)

    def step():
        f = (k_xz @ vt) @ p  # left to right as the reference writes it
        if workload["cost"] == "gaussian":
            dc = (1 / 0.01) * (f - y[:, None])
        elif workload["cost"] == "bernoulli":
            pr = torch.clip(torch.sigmoid(f), 1e-10, 1 - 1e-10)
            dc = -y[:, None] * (1 - pr) + (1 - y[:, None]) * pr
        else:
            dc = -2 * y[:, None] / f + 2 * f
        return ((-eta * vt.T) @ k_zx) @ dc - eta * torch.diag(torch.reciprocal(lam)) @ p + math.sqrt(2 * eta) * xi

    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps * (n_full / n_s)
    return {"value": j / (ms * 1e-3), "unit": UNIT, "ms_per_step_extrapolated": ms,
            "what": (f"reference algebra on this GPU with torch fp64 matmul (cuBLAS DGEMM), dense Gram cached, rows {n_s}/{n_full} at full "
                     f"M={workload['m']}, J={j}, scaled linearly to N (noise pre-drawn on the device)")}


# ---- FP64 peak -----------------------------------------------------------------------------------------------------------
def measure_fp64_peak(n: int = 8192, reps: int = 6) -> float:
    """cuBLAS DGEMM n^3 via torch.matmul, best of `reps` (TFLOP/s): the FP64 denominator MEASURED_PEAKS.json lacks."""
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * n**3 / best * 1e-9


# ---- kernel timing -----------------------------------------------------------------------------------------------------------------
class RoleTimer:
    """CUDA-event pairs around every launch of the contraction kernel, taken INSIDE the library on the launching stream
    (pls_profile_begin / pls_profile_end, include/pls_b200.h): forward / backward ms, launches and algorithmic flops."""

    def __init__(self, ctx):
        self.ctx = ctx

    def begin(self):
        self.ctx.check(self.ctx.lib.pls_profile_begin(self.ctx.handle))

    def end(self) -> dict:
        import ctypes as C

        out = (C.c_double * 6)()
        self.ctx.check(self.ctx.lib.pls_profile_end(self.ctx.handle, out))
        res = {}
        for kind, o in (("forward", 0), ("backward", 3)):
            ms, launches, flops = out[o], int(out[o + 1]), out[o + 2]
            if launches:
                res[kind] = {"launches": launches, "ms_total": ms, "tflops": flops / ms * 1e-9, "flops_per_launch": flops / launches,
                             "ms_per_launch": ms / launches}
        ms, fl = out[0] + out[3], out[2] + out[5]
        res["both"] = {"launches": int(out[1] + out[4]), "ms_total": ms, "tflops": fl / ms * 1e-9 if ms > 0 else 0.0}
        return res


class ReduceTimer:
    """The gradient_reduce hook of a row-sharded run with a CUDA-event pair around every all-reduce."""

    def __init__(self, hook):
        self.hook, self.enabled, self.events, self.bytes = hook, False, [], 0

    def __call__(self, g):
        if not self.enabled:
            return self.hook(g)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.hook(g)
        e1.record()
        self.events.append((e0, e1))
        self.bytes += g.numel() * g.element_size()

    def summary(self, steps: int) -> dict:
        ms = sum(a.elapsed_time(b) for a, b in self.events)
        return {"allreduce_calls": len(self.events), "allreduce_bytes_per_step": self.bytes // max(steps, 1),
                "allreduce_ms_per_step": ms / max(steps, 1)}


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version from C when NCCL_DEBUG is
    set on the box), so file descriptor 1 is pointed at stderr for the whole run and the line goes to a saved duplicate."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


class Runner:
    """One process = one GPU.  Builds a (row shards x particle shards) placement of a workload and times Langevin steps on it."""

    def __init__(self, args):
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist_mod

            self.dist = dist_mod
            self.dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        from projected_langevin_sampling_b200 import _native

        self.ctx = _native.context()
        self.prof = RoleTimer(self.ctx)
        self.seed = 2024
        self._groups = {}

    def max_over_ranks(self, v: float) -> float:
        if self.dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def placement(self, n_groups: int, j_groups: int):
        from projected_langevin_sampling_b200.distributed import GridPlacement, gradient_allreduce, make_row_group

        grid = GridPlacement(rank=self.rank, world=self.world, n_groups=n_groups, j_groups=j_groups)
        key = (n_groups, j_groups)
        if key not in self._groups:  # new_group is collective: every rank creates every grid it will use, in the same order
            self._groups[key] = make_row_group(grid) if self.world > 1 else None
        hook = gradient_allreduce(self._groups[key])
        return grid, (ReduceTimer(hook) if hook is not None else None)

    def build(self, workload, inputs, n_groups, j_groups, gram_mode, j_total=None, local=False, **basis_kw):
        """-> dict(pls, particles, grid, reduce, j_off, j_local, n_local).  j_total = None: the workload's J split over the
        particle shards; otherwise every particle shard owns j_total / j_groups of j_total."""
        x, y, z, ls, outputscale = inputs
        if local:  # this GPU alone holds the whole problem, whatever the world size
            from projected_langevin_sampling_b200.distributed import GridPlacement

            grid, reduce = GridPlacement(rank=0, world=1, n_groups=1, j_groups=1), None
        else:
            grid, reduce = self.placement(n_groups, j_groups)
        r0, r1 = grid.rows(x.shape[0])
        j_off, j_end = grid.particles(j_total if j_total is not None else workload["j"])
        xs, ys = (x, y) if n_groups == 1 else (x[r0:r1].contiguous(), y[r0:r1].contiguous())
        pls = make_pls(workload, xs, ys, z, ls, outputscale, gradient_reduce=reduce, gram_cache=gram_mode, **basis_kw)
        particles = pls.initialise_particles(j_end - j_off, seed=1000 + grid.j_index)  # a row group shares its particles
        pls.basis.engine(j_end - j_off)  # workspaces (and the Gram cache, when used) are setup, like the reference's K_zx
        torch.cuda.synchronize()
        return {"pls": pls, "particles": particles, "grid": grid, "reduce": reduce, "j_off": j_off, "j_local": j_end - j_off,
                "n_local": r1 - r0, "step_no": 0}

    def steps(self, run, eta, count):
        pls, p = run["pls"], run["particles"]
        for _ in range(count):
            pls.step_(p, eta, philox=(self.seed, run["step_no"], run["j_off"]))
            run["step_no"] += 1

    def timed(self, run, eta, steps, clocks=False):
        """barrier + synchronize | exactly `steps` steps between two CUDA events | synchronize + barrier; max over ranks."""
        sampler = ClockSampler(self.local_rank) if clocks else None
        self.barrier()
        if sampler:
            sampler.start()
        launches0 = self.ctx.launches
        if run["reduce"] is not None:
            run["reduce"].enabled, run["reduce"].events, run["reduce"].bytes = True, [], 0
        self.prof.begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.steps(run, eta, steps)
        e1.record()
        torch.cuda.synchronize()
        ksum = self.prof.end()
        self.barrier()
        if run["reduce"] is not None:
            run["reduce"].enabled = False
        ms_local = e0.elapsed_time(e1)
        out = {"ms_total": self.max_over_ranks(ms_local), "ms_total_this_rank": ms_local, "kernels": ksum,
               "launches": self.ctx.launches - launches0}
        if sampler:
            out["clocks"] = sampler.stop()
        return out


def free_run(run):
    run["pls"].basis._engines.clear()
    run.clear()
    torch.cuda.empty_cache()


def small_config_object(runner, name: str, peak: float, steps: int = 50) -> dict:
    """BASELINE configs 2 and 3 (Bernoulli + sigmoid at N=10k, M=64, J=1024; Poisson + square at N=100k, M=256, J=4096) on
    one GPU, default generated Gram: ms/step, particle-updates/s, contraction TFLOP/s and its fraction of the FP64 peak."""
    workload = dict(WORKLOADS[name])
    run = runner.build(workload, synth(workload), 1, 1, False)
    eta = step_size_of(workload)
    runner.steps(run, eta, 5)
    t = runner.timed(run, eta, steps)
    ms = t["ms_total"] / steps
    n, m, j = workload["n"], workload["m"], workload["j"]
    obj = {"workload": workload["label"], "cost": workload["cost"], "value": j / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "step_tflops": 4.0 * n * m * j / (ms * 1e-3) * 1e-12, "kernel_tflops": t["kernels"]["both"]["tflops"],
           "frac": t["kernels"]["both"]["tflops"] / peak if peak else None,
           "per_role": {k: {"tflops": v["tflops"], "ms_per_launch": v["ms_per_launch"]} for k, v in t["kernels"].items() if k != "both"},
           "kernel_share_of_step": t["kernels"]["both"]["ms_total"] / t["ms_total"],
           "particles_finite": bool(torch.isfinite(run["particles"]).all()), "M_k": run["pls"].basis.approximation_dimension,
           "eigenvalue_threshold": workload.get("threshold", 0.0), "step_size": eta}
    free_run(run)
    return obj


def row_sharded_parity(runner, n_groups: int, j_groups: int, steps: int = 3) -> dict:
    """tools/check_row_sharding.py's check on a C5-SHAPED slice (D=16 ARD, M=4096, 65 536 rows, 256 particles per particle
    shard): every rank runs the slice whole on its own GPU and as its cell of the (row x particle) grid with the NCCL gradient
    all-reduce; the particles after `steps` steps must agree (summation order differs: round-off, not bitwise)."""
    w = dict(WORKLOADS["c5"])
    w["n"], w["j"] = 65536, 256 * j_groups
    inputs = synth(w)
    eta = 1e-7
    full = runner.build(w, inputs, 1, 1, False, local=True, eigh_device="cuda", eigenvalue_threshold=1e-6)
    # (a 1 x 1 "grid" of this rank alone; its particles are re-drawn below so that both runs start from the same matrix)
    basis = full["pls"].basis
    m_k = basis.approximation_dimension
    p0 = torch.randn(m_k, w["j"], generator=torch.Generator().manual_seed(5), dtype=torch.float64).cuda()
    p_full = p0.clone()
    for s in range(steps):
        full["pls"].step_(p_full, eta, philox=(77, s, 0))
    eig = (basis.eigenvalues.clone(), basis.eigenvectors.clone())
    free_run(full)
    shard = runner.build(w, inputs, n_groups, j_groups, False, eigendecomposition=eig, eigenvalue_threshold=-1.0)
    j0, j1 = shard["j_off"], shard["j_off"] + shard["j_local"]
    p = p0[:, j0:j1].clone()
    for s in range(steps):
        shard["pls"].step_(p, eta, philox=(77, s, j0))
    err = ((p - p_full[:, j0:j1]).abs().max() / p_full.abs().max()).item()
    moved = ((p_full - p0).abs().max() / p0.abs().max()).item()
    free_run(shard)
    return {"max_rel_err": runner.max_over_ranks(err), "grid": f"{n_groups}x{j_groups}", "rows": w["n"], "D": w["d"], "M": w["m"], "M_k": m_k,
            "J": w["j"], "steps": steps, "relative_change_of_the_particles_over_the_run": moved, "tolerance": 1e-10,
            "what": "C5-shaped slice: particles of the row-sharded run (NCCL all-reduce of the gradient) vs the same slice run whole on one GPU"}


def grid_object(runner, workload, inputs, n_groups, j_groups, steps, warmup, peak, gram_mode=False) -> dict:
    """Informational: the workload on an (n_groups x j_groups) grid -- rows sharded, gradient all-reduced over NCCL per step."""
    t0 = time.perf_counter()
    run = runner.build(workload, inputs, n_groups, j_groups, gram_mode)
    setup_s = time.perf_counter() - t0
    eta = step_size_of(workload)
    runner.steps(run, eta, warmup)
    t = runner.timed(run, eta, steps)
    ms = t["ms_total"] / steps
    n, m, j = workload["n"], workload["m"], workload["j"]
    obj = {"workload": workload["label"], "grid": f"{n_groups}x{j_groups}", "scaling": "strong", "value": j / (ms * 1e-3), "unit": UNIT,
           "ms_per_step": ms, "steps": steps, "warmup": warmup, "rows_per_gpu": run["n_local"], "J_per_gpu": run["j_local"], "J_global": j,
           "step_tflops_per_gpu": 4.0 * run["n_local"] * m * run["j_local"] / (ms * 1e-3) * 1e-12,
           "kernel_tflops_this_gpu": t["kernels"]["both"]["tflops"], "frac": t["kernels"]["both"]["tflops"] / peak if peak else None,
           "particles_finite": bool(torch.isfinite(run["particles"]).all()), "setup_s": setup_s}
    if run["reduce"] is not None:
        obj.update(run["reduce"].summary(steps))
        obj["allreduce_share_of_step"] = obj["allreduce_ms_per_step"] / ms
        obj["collective"] = f"ncclAllReduce(sum, f64) of the M x ld(J_local) gradient over the {n_groups} ranks of a row group, once per step"
    free_run(run)
    return obj


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", type=str, default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational objects measured after the headline")
    ap.add_argument("--grid", type=str, default=None,
                    help="RxC = row shards x particle shards of the ONE named problem (needs R*C == world).  Default 1xN: the J particles "
                         "are split over the N GPUs against replicated data (strong scaling, no communication)")
    ap.add_argument("--gram", type=str, default="generated", choices=["auto", "cached", "staged", "generated"],
                    help="generated (default, the path BASELINE.json's north_star names: Gram tiles recomputed inside the kernels, "
                         "nothing N x M in memory), cached (opt-in: k(X, Z) kept resident in HBM and streamed -- C4: 8.2 GB per GPU) or "
                         "staged (a chunk-sized buffer re-formed every step); auto caches when it fits comfortably")
    ap.add_argument("--rows", dest="n", type=int, default=None, help="override N (debugging; the line then names the reduced workload)")
    ap.add_argument("--particles", dest="j", type=int, default=None)
    args = ap.parse_args()
    workload = dict(WORKLOADS[args.workload])
    if args.n or args.j:
        workload["n"] = args.n or workload["n"]
        workload["j"] = args.j or workload["j"]
        workload["label"] += f" [OVERRIDDEN to N={workload['n']} J={workload['j']}: not the named config]"
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference_arm(args, workload)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the PLS hot path has no CPU fallback (use --impl reference for the CPU arm)")

    runner = Runner(args)
    world, rank, dist = runner.world, runner.rank, runner.dist
    n_groups, j_groups = (int(v) for v in args.grid.lower().split("x")) if args.grid else (1, world)
    grid_name = f"{n_groups}x{j_groups}"
    gram_mode = {"auto": "auto", "cached": True, "staged": "staged", "generated": False}[args.gram]
    t_setup = time.perf_counter()
    inputs = synth(workload)
    run = runner.build(workload, inputs, n_groups, j_groups, gram_mode)
    pls, particles, j_local, j_off, n_local = run["pls"], run["particles"], run["j_local"], run["j_off"], run["n_local"]
    m_k = pls.basis.approximation_dimension
    lam_min = float(pls.basis.eigenvalues.min())
    eta = step_size_of(workload)
    eng0 = pls.basis.engine(j_local)
    gram_note = ("re-formed every step into a chunk-sized staging buffer (%.2f GB; no N x M array)" % (eng0.kstage.numel() * 8 / 1e9)
                 if eng0.kstage is not None else
                 "generated inside the kernels from the points (no N x M array)" if eng0.gram is None else
                 f"cached in HBM ({eng0.gram.numel() * 8 / 1e9:.2f} GB, computed once at setup as the reference's K_zx is) and streamed")
    gram_key = "_cached" if eng0.gram is not None else ("_staged" if eng0.kstage is not None else "")
    chunk_rows, n_chunks, splits = eng0.chunk_rows, len(eng0.chunks), eng0.splits
    del eng0
    setup_s = time.perf_counter() - t_setup

    min_warmup = 1 if args.workload == "c5" else 3  # C5 steps take ~20 s each; its line is labelled accordingly
    warmup = max(args.warmup, min_warmup)
    runner.steps(run, eta, warmup)
    torch.cuda.synchronize()

    # ---- device-resident timed region: exactly `steps` steps -----------------------------------------------------------------------
    t = runner.timed(run, eta, args.steps, clocks=True)
    ms_total, ksum, launches, clocks = t["ms_total"], t["kernels"], t["launches"], t["clocks"]
    finite = bool(torch.isfinite(particles).all())
    ms_per_step = ms_total / args.steps
    j_global = workload["j"]
    value = j_global * args.steps / (ms_total * 1e-3)
    headline_reduce = run["reduce"].summary(args.steps) if run["reduce"] is not None else None

    # ---- end-to-end through the reference-facing API with pinned host buffers ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        p_host = particles.cpu().pin_memory()
        d_host = torch.empty_like(p_host).pin_memory()
        torch.set_default_dtype(torch.float64)  # the reference's noise draw is in the default dtype (README.md:86-87)
        torch.manual_seed(7)
        e2e_steps = max(1, min(args.steps, 5))

        def e2e_step():
            p_dev = p_host.to("cuda", non_blocking=True)  # H2D: particles
            delta = pls.calculate_particle_update(p_dev, eta)  # draws xi on the host generator + H2D, as the reference
            d_host.copy_(delta, non_blocking=True)  # D2H: the step's result
            torch.cuda.synchronize()
            p_host.add_(d_host)

        e2e_step()
        runner.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = runner.max_over_ranks(time.perf_counter() - t0)
        torch.set_default_dtype(torch.float32)
        nbytes = m_k * j_local * 8
        e2e = {"value": j_global * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": 2 * nbytes,
               "d2h_bytes_per_step": nbytes, "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "what": "PLS.calculate_particle_update(particles, step_size) with particles in pinned host memory: H2D particles, "
                       "host torch.normal noise + H2D (the reference's stream), fused step, D2H delta, host add; bytes are per GPU"}
        del p_host, d_host

    # ---- FP64 peak (rank 0 measures, everybody needs it for the informational objects' fractions) ------------------------------------
    peak_live = measure_fp64_peak()
    peak_file = None
    try:
        peak_file = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")))["fp64_tflops_sustained"]
    except Exception:
        pass
    peak = peak_file or peak_live

    # ---- informational objects, after and outside the headline's timed region ------------------------------------------------------
    extras = {}
    do_extras = not args.no_extras and not (args.n or args.j) and args.workload == "c4" and args.grid is None
    if world > 1 and do_extras:
        free_run(run)
        x_steps = max(2, min(args.steps, 5))
        # (1) last round's default: every GPU advances its own J = 4096 particles (weak scaling in J, no communication)
        wk = runner.build(workload, inputs, 1, world, False, j_total=workload["j"] * world)
        runner.steps(wk, eta, 2)
        tw = runner.timed(wk, eta, x_steps)
        ms_w = tw["ms_total"] / x_steps
        extras["weak"] = {"scaling": "weak", "J_per_gpu": wk["j_local"], "J_global": workload["j"] * world, "ms_per_step": ms_w, "steps": x_steps,
                          "value": workload["j"] * world / (ms_w * 1e-3), "unit": UNIT,
                          "kernel_tflops_this_gpu": tw["kernels"]["both"]["tflops"], "frac": tw["kernels"]["both"]["tflops"] / peak,
                          "what": "every GPU advances its own J = 4096 particles against replicated data (round 1's default)"}
        free_run(wk)
        # (2) the same named shape with the ROWS sharded 2 ways as well: NCCL all-reduce of the gradient, timed
        if world % 2 == 0:
            extras["row_sharded"] = grid_object(runner, workload, inputs, 2, world // 2, x_steps, 2, peak)
            extras["row_sharded_parity"] = row_sharded_parity(runner, 2, world // 2)
        # (3) BASELINE config 5 on the whole box
        if world == 8:
            del inputs
            w5 = dict(WORKLOADS["c5"])
            extras["c5"] = grid_object(runner, w5, synth(w5), 2, 4, 2, 1, peak)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline / baselines (rank 0) ---------------------------------------------------------------------------------------------
    traffic, traffic_all, traffic_src = None, {}, None
    try:
        traffic_all = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic_rec = traffic_all.get(args.workload + gram_key)
        if traffic_rec and not (args.n or args.j) and world == 1:
            traffic = traffic_rec["per_launch_bytes_mean"]
            traffic_src = traffic_rec.get("source")
    except Exception:
        pass
    n, m, j = workload["n"], workload["m"], workload["j"]
    achieved = ksum["both"]["tflops"]
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
        "traffic": traffic,
        "traffic_note": ("NOT measured in this run: DRAM bytes per launch (mean of the forward and backward roles) read from profiles/roofline_traffic.json, "
                         "one ncu --set full capture of this command" + (f" ({traffic_src})" if traffic_src else "") + ".  The step is TWO kernels "
                         "with an N_chunk x J round trip between them: the forward writes the cost-derivative chunk Dc to HBM once and the backward "
                         f"reads it once ({chunk_rows * j_local * 8 / 1e9:.1f} GB each way per {chunk_rows}-row chunk here; at C4 ~66 GB per step, ~500x SURVEY 8d's 0.15 GB/step minimum for a single-pass "
                         "design, ~2 % of the HBM bandwidth, not the limiter: the kernels are FP64-pipe bound)"),
        "kernel": "pls::gen_gemm_kernel (forward + backward roles; FP64 DMMA.8x8x4, no tcgen05 kind exists for f64); Gram " + gram_note,
        "algorithmic_flops_per_step": 4.0 * n * m * j,
        "algorithmic_flops_per_step_this_gpu": 4.0 * n_local * m * j_local,
        "per_role": {k: ksum[k] for k in ("forward", "backward") if k in ksum},
        "kernel_share_of_step_this_gpu": ksum["both"]["ms_total"] / t["ms_total_this_rank"],
        "peak_source": ("cuBLAS DGEMM via torch.matmul fp64 8192^3 measured on this pool's B200 "
                        f"(profiles/fp64_peak_r01.json sustained={peak_file}, live best-of-6 in this run={peak_live:.2f}); "
                        "MEASURED_PEAKS.json holds no FP64 figure; FP64 pipe peak from tools/fp64_microbench = 37.1 TFLOP/s"),
        "step_tflops_per_gpu": 4.0 * n_local * m * j_local / (ms_per_step * 1e-3) * 1e-12,
    }
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_arm_subprocess(args.workload, steps=2, warmup=1) if not (args.n or args.j) else None

    if world == 1 and do_extras and not args.no_cpu_baseline:
        try:
            extras["library_bar"] = library_bar_sample(workload, pls, eta)
        except torch.OutOfMemoryError as exc:  # a reported comparison, never a reason to lose the line
            extras["library_bar"] = {"unavailable": str(exc).splitlines()[0]}
        # the same steps with the two opt-in Gram modes -- cached (k(X, Z) computed once and kept in HBM, as the reference keeps
        # K_zx) and staged (k(X_c, Z) re-formed every step into one chunk-sized buffer shared by the chunk's forward and backward
        # launches; nothing N x M kept)
        notes = {"gram_cached": "OrthonormalBasis(gram_cache=True): k(X, Z) kept resident in HBM and loaded by pls_*_cached_f64 instead of "
                                "regenerated; opt-in because north_star specifies on-the-fly Gram tiles",
                 "gram_staged": "OrthonormalBasis(gram_cache='staged'): nothing N x M kept; every step pls_gram_fill_f64 re-forms k(X_c, Z) "
                                "for the row chunk in flight into one chunk-sized buffer that the chunk's forward and backward launches "
                                "stream (instead of each of their column tiles regenerating it)"}
        x_steps = max(2, min(args.steps, 10))
        for key, mode in (("gram_cached", True), ("gram_staged", "staged")):
            try:
                pls.basis._gram_cache_mode, pls.basis._gram = mode, None
                pls.basis._engines.clear()
                q = {"pls": pls, "particles": particles.clone(), "reduce": None, "j_off": j_off, "step_no": 0}
                runner.steps(q, eta, 2)
                tq = runner.timed(q, eta, x_steps)
                ms_c = tq["ms_total"] / x_steps
                ks_c = tq["kernels"]
                eng_c = pls.basis.engine(j_local)
                buf = eng_c.gram if eng_c.gram is not None else eng_c.kstage
                extras[key] = {"value": j_local / (ms_c * 1e-3), "unit": UNIT, "ms_per_step": ms_c, "steps": x_steps,
                               "tflops": ks_c["both"]["tflops"], "frac": ks_c["both"]["tflops"] / peak if peak else None,
                               "step_tflops": 4.0 * n * m * j_local / (ms_c * 1e-3) * 1e-12,
                               "per_role_tflops": {k: ks_c[k]["tflops"] for k in ("forward", "backward") if k in ks_c},
                               "gram_bytes": int(buf.numel() * 8) if buf is not None else 0,
                               "traffic": ((traffic_all.get(args.workload + "_cached") or {}).get("per_launch_bytes_mean")
                                           if key == "gram_cached" else None),
                               "what": "same launch sequence with " + notes[key]}
                del eng_c, buf, q
            except torch.OutOfMemoryError as exc:
                extras[key] = {"unavailable": str(exc).splitlines()[0]}
            finally:
                pls.basis._gram_cache_mode, pls.basis._gram = False, None
                pls.basis._engines.clear()
                torch.cuda.empty_cache()
        if workload["cost"] == "gaussian":
            # the opt-in Gaussian/identity re-association (LangevinEngine._normal_equations) that turns the step into M x M algebra
            # after one N M^2 contraction -- NOT the headline
            try:
                pls.basis._gaussian_normal_equations = True
                pls.basis._engines.clear()
                q = particles.clone()
                t0 = time.perf_counter()
                pls.step_(q, eta, philox=(runner.seed, 0, j_off))
                torch.cuda.synchronize()
                setup_ms = (time.perf_counter() - t0) * 1e3
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for k in range(20):
                    pls.step_(q, eta, philox=(runner.seed, 1 + k, j_off))
                ev1.record()
                torch.cuda.synchronize()
                ms = ev0.elapsed_time(ev1) / 20
                extras["gaussian_normal_equations"] = {
                    "value": j_local / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "setup_ms": setup_ms,
                    "what": "opt-in OrthonormalBasis(gaussian_normal_equations=True): k(Z,X)k(X,Z)/s and k(Z,X)y/s formed once "
                            "(setup_ms, 2 N M^2 flops), then 2 M^2 J flops per step; same particles to round-off "
                            "(tests/test_gpu_configs.py::test_gaussian_normal_equations_shortcut); Gaussian cost only"}
            except torch.OutOfMemoryError as exc:
                extras["gaussian_normal_equations"] = {"unavailable": str(exc).splitlines()[0]}
            finally:
                pls.basis._gaussian_normal_equations = False
                pls.basis._engines.clear()
        free_run(run)
        del pls, particles
        # BASELINE configs 2 and 3: the non-Gaussian cost epilogues, driver-timed
        for name in ("c2", "c3"):
            try:
                extras[name] = small_config_object(runner, name, peak)
            except Exception as exc:
                extras[name] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}

    config = base_config(workload, world, grid_name)
    config.update({
        "M_k": m_k, "J_per_gpu": j_local, "rows_per_gpu": n_local, "lambda_min": lam_min, "gram": gram_note,
        "row_chunk": {"rows": chunk_rows, "chunks_per_step": n_chunks, "backward_splits": splits},
        "noise": "Philox4x32-10 on device keyed on global (row, particle)",
        "parallelism": ((f"grid {grid_name}: rows sharded x{n_groups} (NCCL all-reduce of the M x J_local gradient per step), " if n_groups > 1 else "")
                        + f"particles sharded x{j_groups} (J_local = {j_local} of J = {j_global}; no per-step communication)"),
        "step_size_note": "SURVEY 8d names eta = 1e-6 for C4; lambda_min = 3.6e-9 makes the prior term eta/lambda diverge there, so 1e-9 is used "
                          "(the cost of a step does not depend on eta)" if workload["cost"] == "gaussian" else None,
        "particles_finite": finite, "setup_s": setup_s})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
    }
    if headline_reduce is not None:
        line["allreduce"] = headline_reduce
    line.update(extras)
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
