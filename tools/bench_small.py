"""Eager loop vs CUDA-graph replay of the Langevin step at the small BASELINE configs (launch-bound regime).

    python tools/bench_small.py --workload c2 --steps 200
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"])
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    w = dict(bench.WORKLOADS[args.workload])
    x, y, z, ls, os_ = bench.synth(w)
    pls = bench.make_pls(w, x, y, z, ls, os_)
    p = pls.initialise_particles(w["j"], seed=1)
    eta = 1e-6
    out = {"workload": w["label"], "steps": args.steps}
    def timed(q, steps, graph):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pls.run(q, eta, steps, seed=1, cuda_graph=graph)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    for mode in ("eager", "cuda_graph"):
        graph = mode == "cuda_graph"
        q = p.clone()
        timed(q, 5, graph)
        # marginal cost of a step: a graph-mode call also pays one capture + instantiation, which the difference removes
        short, long_ = timed(q, args.steps // 10, graph), timed(q, args.steps + args.steps // 10, graph)
        dt = long_ - short
        out[mode] = {"us_per_step": round(dt / args.steps * 1e6, 1), "particle_updates_per_s": round(w["j"] * args.steps / dt, 1),
                     "step_tflops": round(4.0 * w["n"] * w["m"] * w["j"] * args.steps / dt * 1e-12, 2),
                     "call_overhead_ms": round((short - dt / args.steps * (args.steps // 10)) * 1e3, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
