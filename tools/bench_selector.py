"""Times the ConditionalVariance selector (pls_cv_select_f64) at a given shape against its HBM roofline.

    python tools/bench_selector.py --n 1000000 --d 8 --m 1024

Algorithmic bytes: 8*N*sum_{i<M-1} i (rows of C streamed once per pivot) + per pivot the point
set (N*SP*8), d read+write (16 N), the new row of C (8 N).
"""
import argparse
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_langevin_sampling_b200 import _native as nat, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=8)
    ap.add_argument("--m", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=1)
    args = ap.parse_args()
    ctx = nat.context()
    n, d, m = args.n, args.d, args.m
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, d, generator=g, dtype=torch.float64).cuda()
    inv_ls = [1.0 / (math.sqrt(d) * (0.75 + 0.5 * k / max(d - 1, 1))) for k in range(d)]
    xa = ops.prepare_points(ctx, nat.KERNEL_RBF, x, inv_ls, x.mean(0).tolist(), 0.0)
    sp = xa.shape[1]
    best = 1e30
    for _ in range(args.reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        idx, nsel = ops.cv_select(ctx, nat.KERNEL_RBF, xa, d, 1.0, m, 1e-12, 0.0)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    stream_bytes = 8.0 * n * (m - 1) * (m - 2) / 2
    per_pivot = (m - 1) * n * (sp * 8 + 16 + 8 + 1)
    peak = 6540.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass
    gbs = (stream_bytes + per_pivot) / best / 1e9
    print(json.dumps({"n": n, "d": d, "m": m, "selected": nsel, "seconds": round(best, 4), "algorithmic_GB": round((stream_bytes + per_pivot) / 1e9, 2),
                      "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3), "first_indices": idx[:5].tolist()}))


if __name__ == "__main__":
    main()
