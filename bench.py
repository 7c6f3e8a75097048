#!/usr/bin/env python
"""bench.py -- particle-updates/sec of the fused PLS Langevin step (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c4|c3|c2|c5] [--grid RxC] [--gram generated|staged|cached|auto]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one Langevin step over all J particles of the workload:
  C4 (default, the configuration the metric is quoted on): N=1,000,000 D=8 ARD, M=1024, J=4096, Gaussian cost.
At --gpus N the particles are sharded (weak scaling: every GPU advances its own J particles against replicated
X, y, Z, V~, lambda; no per-step communication; Philox noise keyed on the global particle index).

The JSON line carries: value (device-timed, inputs resident in HBM), e2e (through the reference-facing API with pinned
HOST buffers, copies inside the timed region), roofline (FP64 tensor, live CUDA-event kernel timing), cpu_baseline
(the oracle's reference-style torch-CPU step on the box's host cores, bounded sample), clocks, gpu_launches.
Every timed number above is the DEFAULT path (Gram tiles regenerated inside the kernels, nothing N x M in memory, --gram generated).
At N = 1 the line also carries informational measurements taken after and outside the timed region: library_bar (the
reference's algebra on cuBLAS on this GPU), gram_cached / gram_staged (the same steps with the two opt-in Gram modes) and
gaussian_normal_equations (the opt-in M x M re-association for the Gaussian cost).

`--impl reference` times the reference's own CPU formulation of the path (the oracle port -- the reference itself needs
gpytorch, which is not installed) on all host threads and prints the same line shape with "impl": "reference".
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

import torch  # noqa: E402

METRIC = "PLS particle-updates/sec at N=1M,M=1024,J=4096; % of FP64/HBM roofline"
UNIT = "particle-updates/s"

WORKLOADS = {
    # name: (N, D, M, J, cost)
    "c4": dict(n=1_000_000, d=8, m=1024, j=4096, cost="gaussian", label="C4 UCI-scale synthetic regression N=1M D=8 ARD M=1024 J=4096"),
    "c5": dict(n=20_000_000, d=16, m=4096, j=16384, cost="gaussian",
               label="C5 large synthetic regression N=20M D=16 ARD M=4096 J=16384 (row- and particle-sharded, NCCL gradient all-reduce)"),
    "c3": dict(n=100_000, d=1, m=256, j=4096, cost="poisson", label="C3 Poisson f^2 regression N=100k D=1 M=256 J=4096"),
    "c2": dict(n=10_000, d=1, m=64, j=1024, cost="bernoulli", label="C2 1D Bernoulli classification N=10k M=64 J=1024"),
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


def make_pls(workload: dict, x, y, z, ls, outputscale, gradient_reduce=None, gram_cache="auto"):
    import projected_langevin_sampling_b200 as pkg
    from projected_langevin_sampling_b200.projected_langevin_sampling import costs, link_functions as lf

    kernel = pkg.ScaleKernel(pkg.RBFKernel(ard_num_dims=x.shape[1], lengthscale=ls), outputscale=outputscale)
    basis = pkg.OrthonormalBasis(pkg.PLSKernel(kernel, z), z, x, eigenvalue_threshold=workload.get("threshold", 0.0), verbose=False,
                                 gradient_reduce=gradient_reduce, gram_cache=gram_cache)
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


# ---- CPU baseline (oracle port of the reference's CPU path) -----------------------------------------------------------
def cpu_reference_sample(workload: dict, steps: int, warmup: int):
    """The reference-style torch-CPU Langevin step (oracle.pls_oracle.reference_style_cpu_step: left-to-right dense
    matmuls with the N x M Gram cached, eigh(eye) + torch.normal per step) on a bounded sample: a 1/20 row slice at full
    M with J_c = 256 particles.  The J-independent products (k(X,Z) V~ and V~^T k(Z,X), which the reference re-forms
    every step) and the J-dependent rest are timed separately and scaled linearly to the full N and J."""
    from oracle.pls_oracle import OrthonormalBasisOracle, RBFScaleKernel, langevin_noise

    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_default_dtype(torch.float64)
    x, y, z, ls, outputscale = synth(workload)
    n_full, j_full = workload["n"], workload["j"]
    n_s = max(min(n_full, 1000), n_full // 20)
    j_c = min(256, j_full)
    xs, ys = x[:n_s], y[:n_s]
    basis = OrthonormalBasisOracle(RBFScaleKernel(ls, outputscale), z, xs)
    k_zx, vt, lam = basis.k_zx.contiguous(), basis.scaled_eigenvectors, basis.eigenvalues
    p = torch.randn(vt.shape[1], j_c, generator=torch.Generator().manual_seed(1))
    eta = 1e-9
    t_fixed, t_rest = [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        phi = k_zx.T @ vt  # (N_s, M_k): re-formed by the reference every step (orthonormal.py:106-108)
        back = (-eta * vt.T) @ k_zx  # (M_k, N_s): likewise (orthonormal.py:151-154)
        t1 = time.perf_counter()
        f = phi @ p
        if workload["cost"] == "gaussian":
            dc = (1 / 0.01) * (f - ys[:, None])
        elif workload["cost"] == "bernoulli":
            pr = torch.clip(torch.reciprocal(1 + torch.exp(-f)), 1e-10, 1 - 1e-10)
            dc = -ys[:, None] * (1 - pr) + (1 - ys[:, None]) * pr
        else:
            dc = -2 * ys[:, None] / f + 2 * f
        xi = langevin_noise(p.shape[0], j_c)  # eigh(eye(M_k)) + torch.normal, as samplers.py:27-35
        delta = back @ dc - eta * torch.diag(torch.reciprocal(lam)) @ p + math.sqrt(2 * eta) * xi
        p = p + delta
        t2 = time.perf_counter()
        if it >= warmup:
            t_fixed.append(t1 - t0)
            t_rest.append(t2 - t1)
    scale_n = n_full / n_s
    tf, tr = statistics.median(t_fixed), statistics.median(t_rest)
    t_step_full = scale_n * (tf + tr * (j_full / j_c))
    return {
        "value": j_full / t_step_full,
        "unit": UNIT,
        "cores": torch.get_num_threads(),
        "kind": "port",
        "sample": (f"oracle port of the reference torch-CPU step, dense Gram cached, rows {n_s}/{n_full} at full M={workload['m']}, "
                   f"J_c={j_c}: J-independent products {tf:.3f}s + J-dependent {tr:.3f}s per sampled step, scaled linearly to "
                   f"N={n_full}, J={j_full} ({steps} timed steps, median)"),
        "sample_seconds": sum(t_fixed) + sum(t_rest),
        "ms_per_step_extrapolated": t_step_full * 1e3,
    }


def run_reference_arm(args, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    res = cpu_reference_sample(workload, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step_extrapolated"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload["label"], "note": "CPU arm: one host, all threads; value extrapolated from the bounded sample"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
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


# ---- kernel timing hooks ----------------------------------------------------------------------------------------------------
class KernelTimer:
    """CUDA-event pairs around every forward / backward launch of the generated-operand GEMM (on torch's current stream,
    which is the stream the library launches on)."""

    def __init__(self):
        self.records = []  # (kind, flops, e0, e1)
        self.enabled = False

    def install(self):
        from projected_langevin_sampling_b200 import ops

        timer = self
        orig_fwd, orig_bwd = ops.forward, ops.backward

        def fwd(ctx, kernel_id, xa, za, d, w, j, epilogue, out, **kw):
            if not timer.enabled:
                return orig_fwd(ctx, kernel_id, xa, za, d, w, j, epilogue, out, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig_fwd(ctx, kernel_id, xa, za, d, w, j, epilogue, out, **kw)
            e1.record()
            timer.records.append(("forward", 2.0 * xa.shape[0] * za.shape[0] * j, e0, e1))
            return r

        def bwd(ctx, kernel_id, za, xa, d, dc, j, gp, splits, accumulate, **kw):
            if not timer.enabled:
                return orig_bwd(ctx, kernel_id, za, xa, d, dc, j, gp, splits, accumulate, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig_bwd(ctx, kernel_id, za, xa, d, dc, j, gp, splits, accumulate, **kw)
            e1.record()
            timer.records.append(("backward", 2.0 * xa.shape[0] * za.shape[0] * j, e0, e1))
            return r

        ops.forward, ops.backward = fwd, bwd

    def summary(self):
        out = {}
        for kind in ("forward", "backward"):
            recs = [r for r in self.records if r[0] == kind]
            if recs:
                ms = sum(r[2].elapsed_time(r[3]) for r in recs)
                fl = sum(r[1] for r in recs)
                out[kind] = {"launches": len(recs), "ms_total": ms, "tflops": fl / ms * 1e-9, "flops_per_launch": fl / len(recs),
                             "ms_per_launch": ms / len(recs)}
        recs = self.records
        ms = sum(r[2].elapsed_time(r[3]) for r in recs)
        fl = sum(r[1] for r in recs)
        out["both"] = {"launches": len(recs), "ms_total": ms, "tflops": fl / ms * 1e-9 if ms > 0 else 0.0}
        return out


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
    ap.add_argument("--grid", type=str, default=None,
                    help="RxC = row shards x particle shards of ONE problem (strong scaling; needs R*C == world). Default: every GPU "
                         "advances its own J particles against replicated data (weak scaling in J, no communication)")
    ap.add_argument("--gram", type=str, default="generated", choices=["auto", "cached", "staged", "generated"],
                    help="generated (default, the path BASELINE.json's north_star names: Gram tiles recomputed inside the kernels, "
                         "nothing N x M in memory) or cached (opt-in: k(X, Z) kept resident in HBM and streamed -- C4: 8.2 GB per "
                         "GPU); auto caches when it fits comfortably (engine.want_gram_cache).  At N = 1 the default run also "
                         "reports the cached mode in the line's `gram_cached` object")
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

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the PLS hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from projected_langevin_sampling_b200 import _native

    ctx = _native.context()
    t_setup = time.perf_counter()
    x, y, z, ls, outputscale = synth(workload)
    gram_mode = {"auto": "auto", "cached": True, "staged": "staged", "generated": False}[args.gram]
    grid = None
    if args.grid:
        from projected_langevin_sampling_b200.distributed import GridPlacement, gradient_allreduce, make_row_group

        n_groups, j_groups = (int(v) for v in args.grid.lower().split("x"))
        grid = GridPlacement(rank=rank, world=world, n_groups=n_groups, j_groups=j_groups)
        row_group = make_row_group(grid) if world > 1 else None
        r0, r1 = grid.rows(workload["n"])
        j_off, j_end = grid.particles(workload["j"])
        pls = make_pls(workload, x[r0:r1].contiguous(), y[r0:r1].contiguous(), z, ls, outputscale, gradient_allreduce(row_group),
                       gram_cache=gram_mode)
        j_local = j_end - j_off
        del x, y
    else:
        pls = make_pls(workload, x, y, z, ls, outputscale, gram_cache=gram_mode)
        j_local = workload["j"]  # weak scaling: every GPU owns J particles
        j_off = rank * j_local
    m_k = pls.basis.approximation_dimension
    lam_min = float(pls.basis.eigenvalues.min())
    eta = 1e-9 if workload["cost"] == "gaussian" else 1e-6
    particles = pls.initialise_particles(j_local, seed=1000 + (grid.j_index if grid is not None else rank))  # a row group shares its particles
    eng0 = pls.basis.engine(j_local)  # workspaces (and the Gram cache, when used) are setup, like the reference's K_zx
    gram_note = ("re-formed every step into a chunk-sized staging buffer (%.2f GB; no N x M array)" % (eng0.kstage.numel() * 8 / 1e9)
                 if eng0.kstage is not None else
                 "generated inside the kernels from the points (no N x M array)" if eng0.gram is None else
                 f"cached in HBM ({eng0.gram.numel() * 8 / 1e9:.2f} GB, computed once at setup as the reference's K_zx is) and streamed")
    gram_is_cached = eng0.gram is not None
    gram_key = "_cached" if gram_is_cached else ("_staged" if eng0.kstage is not None else "")
    del eng0
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    timer = KernelTimer()
    timer.install()
    seed = 2024
    step_no = 0
    min_warmup = 1 if args.workload == "c5" else 3  # C5 steps take ~20 s each; its line is labelled accordingly
    for _ in range(max(args.warmup, min_warmup)):
        pls.step_(particles, eta, philox=(seed, step_no, j_off))
        step_no += 1
    torch.cuda.synchronize()

    # ---- device-resident timed region -------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = ctx.launches
    timer.enabled = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        pls.step_(particles, eta, philox=(seed, step_no, j_off))
        step_no += 1
    e1.record()
    torch.cuda.synchronize()
    timer.enabled = False
    if dist is not None:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    clocks = sampler.stop()
    finite = bool(torch.isfinite(particles).all())
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    j_global = workload["j"] if grid is not None else workload["j"] * world
    value = j_global * args.steps / (ms_total * 1e-3)
    ksum = timer.summary()

    # ---- end-to-end through the reference-facing API with pinned host buffers ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        p_host = particles.cpu().pin_memory()
        d_host = torch.empty_like(p_host).pin_memory()
        torch.set_default_dtype(torch.float64)  # the reference's noise draw is in the default dtype (README.md:86-87)
        torch.manual_seed(7)
        e2e_steps = max(1, min(args.steps, 3))

        def e2e_step():
            p_dev = p_host.to("cuda", non_blocking=True)  # H2D: particles
            delta = pls.calculate_particle_update(p_dev, eta)  # draws xi on the host generator + H2D, as the reference
            d_host.copy_(delta, non_blocking=True)  # D2H: the step's result
            torch.cuda.synchronize()
            p_host.add_(d_host)

        e2e_step()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        torch.set_default_dtype(torch.float32)
        nbytes = m_k * j_local * 8
        e2e = {"value": j_global * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": 2 * nbytes,
               "d2h_bytes_per_step": nbytes, "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "what": "PLS.calculate_particle_update(particles, step_size) with particles in pinned host memory: H2D particles, "
                       "host torch.normal noise + H2D (the reference's stream), fused step, D2H delta, host add"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline / baselines (rank 0) -----------------------------------------------------------------------------------------
    peak_live = measure_fp64_peak()
    peak_file = None
    try:
        peak_file = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")))["fp64_tflops_sustained"]
    except Exception:
        pass
    peak = peak_file or peak_live
    traffic = None
    traffic_all = {}
    try:
        traffic_all = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic_rec = traffic_all.get(args.workload + gram_key)
        if traffic_rec and not (args.n or args.j):
            traffic = traffic_rec["per_launch_bytes_mean"]
    except Exception:
        pass
    n, m, j = workload["n"], workload["m"], workload["j"]
    achieved = ksum["both"]["tflops"]
    n_local = (grid.rows(n)[1] - grid.rows(n)[0]) if grid is not None else n
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
        "traffic": traffic,
        "traffic_note": "DRAM bytes per launch (mean of the forward and backward roles) from one ncu --set full capture of this command, "
                        "profiles/roofline_traffic.json; the kernel is FP64-pipe bound, traffic ~= the Dc chunk written / read once (+ the chunk's "
                        "rows of the cached Gram read once)",
        "kernel": "pls::gen_gemm_kernel (forward + backward roles; FP64 DMMA.8x8x4, no tcgen05 kind exists for f64); Gram " + gram_note,
        "algorithmic_flops_per_step": 4.0 * n * m * j,
        "algorithmic_flops_per_step_this_gpu": 4.0 * n_local * m * j_local,
        "per_role": {k: ksum[k] for k in ("forward", "backward") if k in ksum},
        "kernel_share_of_step": ksum["both"]["ms_total"] / ms_total if world == 1 else None,
        "peak_source": ("cuBLAS DGEMM via torch.matmul fp64 8192^3 measured on this pool's B200 "
                        f"(profiles/fp64_peak_r01.json sustained={peak_file}, live best-of-6 in this run={peak_live:.2f}); "
                        "MEASURED_PEAKS.json holds no FP64 figure; FP64 pipe peak from tools/fp64_microbench = 37.1 TFLOP/s"),
        "step_tflops_per_gpu": 4.0 * n_local * m * j_local / (ms_per_step * 1e-3) * 1e-12,
    }
    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_reference_sample(workload, steps=2, warmup=1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    library_bar = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            library_bar = library_bar_sample(workload, pls, eta)
        except torch.OutOfMemoryError as exc:  # a reported comparison, never a reason to lose the line
            library_bar = {"unavailable": str(exc).splitlines()[0]}
    extras = {"gram_cached": None, "gram_staged": None}
    if world == 1 and gram_mode is False and not args.no_cpu_baseline:
        # informational measurements, after and outside the headline's timed region: the same steps with the two opt-in Gram
        # modes -- cached (k(X, Z) computed once and kept in HBM, as the reference keeps K_zx) and staged (k(X_c, Z) re-formed
        # every step into one chunk-sized buffer shared by the chunk's forward and backward launches; nothing N x M kept)
        notes = {"gram_cached": "OrthonormalBasis(gram_cache=True): k(X, Z) kept resident in HBM and loaded by pls_*_cached_f64 instead of "
                                "regenerated; opt-in because north_star specifies on-the-fly Gram tiles",
                 "gram_staged": "OrthonormalBasis(gram_cache='staged'): nothing N x M kept; every step pls_gram_fill_f64 re-forms k(X_c, Z) "
                                "for the row chunk in flight into one chunk-sized buffer that the chunk's forward and backward launches "
                                "stream (instead of each of their J/256 column tiles regenerating it)"}
        for key, mode in (("gram_cached", True), ("gram_staged", "staged")):
            try:
                pls.basis._gram_cache_mode, pls.basis._gram = mode, None
                pls.basis._engines.clear()
                q = particles.clone()
                for k in range(2):
                    pls.step_(q, eta, philox=(seed, k, j_off))
                torch.cuda.synchronize()
                timer.records = []
                timer.enabled = True
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for k in range(args.steps):
                    pls.step_(q, eta, philox=(seed, 2 + k, j_off))
                ev1.record()
                torch.cuda.synchronize()
                timer.enabled = False
                ms_c = ev0.elapsed_time(ev1) / args.steps
                ks_c = timer.summary()
                eng_c = pls.basis.engine(j_local)
                buf = eng_c.gram if eng_c.gram is not None else eng_c.kstage
                extras[key] = {"value": j_local / (ms_c * 1e-3), "unit": UNIT, "ms_per_step": ms_c, "steps": args.steps,
                               "tflops": ks_c["both"]["tflops"], "frac": ks_c["both"]["tflops"] / peak if peak else None,
                               "step_tflops": 4.0 * n * m * j_local / (ms_c * 1e-3) * 1e-12,
                               "per_role_tflops": {k: ks_c[k]["tflops"] for k in ("forward", "backward") if k in ks_c},
                               "gram_bytes": int(buf.numel() * 8) if buf is not None else 0,
                               "traffic": ((traffic_all.get(args.workload + "_cached") or {}).get("per_launch_bytes_mean")
                                           if key == "gram_cached" and not (args.n or args.j) else None),
                               "what": "same launch sequence with " + notes[key]}
                del eng_c, buf
            except torch.OutOfMemoryError as exc:
                extras[key] = {"unavailable": str(exc).splitlines()[0]}
            finally:
                pls.basis._gram_cache_mode, pls.basis._gram = False, None
                pls.basis._engines.clear()
                torch.cuda.empty_cache()
    shortcut = None
    if world == 1 and workload["cost"] == "gaussian" and not args.no_cpu_baseline:
        # informational, outside the timed region and NOT the headline: the opt-in Gaussian/identity re-association
        # (LangevinEngine._normal_equations) that turns the step into M x M algebra after one N M^2 contraction
        try:
            pls.basis._gaussian_normal_equations = True
            pls.basis._engines.clear()
            q = particles.clone()
            t0 = time.perf_counter()
            pls.step_(q, eta, philox=(seed, 0, j_off))
            torch.cuda.synchronize()
            setup_ms = (time.perf_counter() - t0) * 1e3
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for k in range(20):
                pls.step_(q, eta, philox=(seed, 1 + k, j_off))
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / 20
            shortcut = {"value": j_local / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "setup_ms": setup_ms,
                        "what": "opt-in OrthonormalBasis(gaussian_normal_equations=True): k(Z,X)k(X,Z)/s and k(Z,X)y/s formed once "
                                "(setup_ms, 2 N M^2 flops), then 2 M^2 J flops per step; same particles to round-off "
                                "(tests/test_gpu_configs.py::test_gaussian_normal_equations_shortcut); Gaussian cost only"}
        except torch.OutOfMemoryError as exc:
            shortcut = {"unavailable": str(exc).splitlines()[0]}
        finally:
            pls.basis._gaussian_normal_equations = False
            pls.basis._engines.clear()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, min_warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if grid is not None else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload["label"], "N": n, "D": workload["d"], "M": m, "M_k": m_k, "J_per_gpu": j_local,
                   "J_global": j_global, "rows_per_gpu": n_local, "cost": workload["cost"], "step_size": eta, "lambda_min": lam_min, "gram": gram_note,
                   "noise": "Philox4x32-10 on device keyed on global (row, particle)", "parallelism": (f"grid {args.grid}: rows sharded x{grid.n_groups} (NCCL all-reduce of the M x J_local gradient per step), "
                                   f"particles sharded x{grid.j_groups}") if grid is not None else f"particle-sharded x{world}",
                   "l2": "per-step working set (Dc chunk 8 GiB written+read) exceeds the 126 MB L2; no flush needed",
                   "particles_finite": finite, "setup_s": setup_s},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "library_bar": library_bar, "gram_cached": extras["gram_cached"], "gram_staged": extras["gram_staged"],
        "gaussian_normal_equations": shortcut,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
