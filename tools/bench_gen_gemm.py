"""Times pls_forward_f64 / pls_backward_f64 alone (CUDA events, after warm-up) at a given shape.

    python tools/bench_gen_gemm.py --n 262144 --m 1024 --d 8 --j 4096 [--epilogue 0|1|2] [--rt 0|1|2] [--reps 3]

Development aid for DESIGN.md's tile-shape table; bench.py is the judged benchmark.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_langevin_sampling_b200 import _native as nat, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=262144)
    ap.add_argument("--m", type=int, default=1024)
    ap.add_argument("--d", type=int, default=8)
    ap.add_argument("--j", type=int, default=4096)
    ap.add_argument("--epilogue", type=int, default=nat.EPI_COST_DERIVATIVE)
    ap.add_argument("--rt", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--splits", type=int, default=0)
    ap.add_argument("--kernel", type=str, default="rbf", choices=["rbf", "linear"])
    ap.add_argument("--cost", type=str, default="gaussian", choices=["gaussian", "poisson", "bernoulli", "student_t", "multimodal"])
    ap.add_argument("--roles", type=str, default="forward,backward")
    ap.add_argument("--cached", action="store_true", help="stream a resident Gram (pls_*_cached_f64) instead of generating it")
    args = ap.parse_args()

    kid = nat.KERNEL_RBF if args.kernel == "rbf" else nat.KERNEL_LINEAR
    ctx = nat.context()
    ctx.lib.pls_set_tile_shape(ctx.handle, args.rt)
    g = torch.Generator().manual_seed(0)
    n, m, d, j = args.n, args.m, args.d, args.j
    x = torch.randn(n, d, generator=g, dtype=torch.float64).cuda()
    z = x[:m].clone()
    inv_ls = [1.0 / (d ** 0.5 * (0.75 + 0.5 * k / max(d - 1, 1))) for k in range(d)]
    centre = z.mean(0).tolist()
    xa = ops.prepare_points(ctx, kid, x, inv_ls, centre, 0.0)
    za = ops.prepare_points(ctx, kid, z, inv_ls, centre, 0.0)
    w = (torch.randn(m, j, generator=g, dtype=torch.float64) / m).cuda()
    y = torch.randn(n, generator=g, dtype=torch.float64).cuda()
    cost = nat.PlsCost()
    cost.cost_id, cost.link_id, cost.closed_form, cost.observation_noise = nat.COST_GAUSSIAN, nat.LINK_IDENTITY, 1, 0.01
    cost.link_jitter, cost.probit_divisor, cost.degrees_of_freedom, cost.scale = 1e-10, 2.0 ** 0.5, 4.0, 0.7
    cost.shift, cost.bernoulli_noise, cost.log_weight_1, cost.log_weight_2, cost.log_normaliser = 1.0, 0.3, -1.2, -0.36, 0.0
    if args.cost == "poisson":
        cost.cost_id, cost.link_id = nat.COST_POISSON, nat.LINK_SQUARE
        y = torch.poisson(torch.full((n,), 3.0, dtype=torch.float64), generator=g).cuda()
    elif args.cost == "bernoulli":
        cost.cost_id, cost.link_id = nat.COST_BERNOULLI, nat.LINK_SIGMOID
        y = (torch.rand(n, generator=g, dtype=torch.float64) > 0.5).double().cuda()
    elif args.cost == "student_t":
        cost.cost_id = nat.COST_STUDENT_T
    elif args.cost == "multimodal":
        cost.cost_id, cost.closed_form, cost.observation_noise = nat.COST_MULTIMODAL, 0, 0.5
    tile_rows = ops.forward_tile_rows(ctx, j)
    out_rows = (n + tile_rows - 1) // tile_rows if args.epilogue == nat.EPI_COST else n
    out = torch.zeros(out_rows, j, dtype=torch.float64).cuda()
    dc = torch.randn(n, j, generator=torch.Generator(device="cuda").manual_seed(1), dtype=torch.float64, device="cuda")
    splits = args.splits or ops.backward_splits(ctx, n, m, j)
    gp = torch.zeros(splits, m, j, dtype=torch.float64).cuda()
    flops = 2.0 * n * m * j
    gram = ops.gram_cache(ctx, kid, xa, za, d) if args.cached else None
    res = {"cached": bool(args.cached), "n": n, "m": m, "d": d, "j": j, "rt": args.rt, "epilogue": args.epilogue, "kernel": args.kernel, "cost": args.cost, "splits": splits}

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    if "forward" in args.roles:
        if args.epilogue == nat.EPI_COST_DERIVATIVE_AND_COST:
            part = torch.zeros((n + tile_rows - 1) // tile_rows, j, dtype=torch.float64).cuda()
            ms = timed(lambda: ops.forward_step(ctx, kid, xa, za, d, w, j, cost, y, out, part, gram=gram))
        else:
            ms = timed(lambda: ops.forward(ctx, kid, xa, za, d, w, j, args.epilogue, out, cost=cost, y=y, gram=gram))
        res["forward_ms"], res["forward_tflops"] = round(ms, 3), round(flops / ms / 1e9, 3)
    if "backward" in args.roles:
        ms = timed(lambda: ops.backward(ctx, kid, za, xa, d, dc, j, gp, splits, accumulate=False, gram=gram))
        res["backward_ms"], res["backward_tflops"] = round(ms, 3), round(flops / ms / 1e9, 3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
