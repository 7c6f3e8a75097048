"""Per-epoch cost of the training loop at the headline shape: fused train_pls (energy from the step's own forward) vs the
reference-shaped loop (calculate_particle_update + calculate_energy_potential = two forwards per epoch).

    python tools/bench_train.py [--workload c4] [--epochs 4]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from projected_langevin_sampling_b200.trainers import train_pls  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4", choices=["c2", "c3", "c4"])
    ap.add_argument("--epochs", type=int, default=4)
    args = ap.parse_args()
    w = dict(bench.WORKLOADS[args.workload])
    x, y, z, ls, os_ = bench.synth(w)
    pls = bench.make_pls(w, x, y, z, ls, os_)
    eta = 1e-9 if w["cost"] == "gaussian" else 1e-6
    p = pls.initialise_particles(w["j"], seed=1)
    out = {"workload": w["label"], "epochs": args.epochs}

    def timed(fn):
        fn(1)  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn(args.epochs)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / args.epochs * 1e3

    def steps_only(n):
        q = p.clone()
        for s in range(n):
            pls.step_(q, eta, philox=(1, s, 0))

    def reference_shaped(n):
        q = p.clone()
        for s in range(n):
            pls.step_(q, eta, philox=(1, s, 0))
            pls.calculate_energy_potential(q)

    def fused(n):
        train_pls(pls, p.clone(), n, eta, early_stopper_patience=1e9, philox_seed=1)

    out["ms_per_step_no_energy"] = round(timed(steps_only), 2)
    out["ms_per_epoch_two_forwards"] = round(timed(reference_shaped), 2)
    out["ms_per_epoch_train_pls_fused"] = round(timed(fused), 2)  # includes the one extra forward after the last epoch, amortised
    print(json.dumps(out))


if __name__ == "__main__":
    main()
