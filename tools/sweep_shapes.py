"""Randomised parity sweep of the contraction kernels (both roles, all forward epilogues, generated and cached Gram) against
dense float64 torch algebra, over many seeded shapes including degenerate ones.  tests/test_gpu_kernels.py holds a fixed
subset; this is the long version for a GPU box:

    python tools/sweep_shapes.py --count 80 --seed 7
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_langevin_sampling_b200 import _native as nat, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=60)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    ctx = nat.context()
    rng = np.random.default_rng(args.seed)
    worst = 0.0
    for it in range(args.count):
        small = it % 3 == 0
        n = int(rng.integers(1, 90 if small else 3000))
        m = int(rng.integers(1, 40 if small else 600))
        d = int(rng.integers(1, 27))  # MAX_D = 26
        j = int(rng.integers(1, 20 if small else (2100 if it % 5 == 1 else 900)))
        ld = j + (j & 1) + int(rng.choice([0, 2, 14, 30]))
        rt = int(rng.choice([0, 1, 2]))
        ctx.lib.pls_set_tile_shape(ctx.handle, rt)
        ctx.lib.pls_set_tile_sets(ctx.handle, int(rng.choice([0, 1, 2])))  # 64 x 512 tiles (tensor-memory parking): rule / never / forced
        g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
        x = torch.randn(n, d, generator=g, dtype=torch.float64).cuda()
        z = torch.randn(m, d, generator=g, dtype=torch.float64).cuda()
        inv_ls = [1.0 / (2.0 + d ** 0.5 + 0.05 * k) for k in range(d)]
        centre = z.mean(0).tolist()
        xa = ops.prepare_points(ctx, nat.KERNEL_RBF, x, inv_ls, centre, 0.0)
        za = ops.prepare_points(ctx, nat.KERNEL_RBF, z, inv_ls, centre, float(np.log(1.4)))
        k_xz = ops.gram(ctx, nat.KERNEL_RBF, xa, za, d)
        w = torch.randn(m, ld, generator=g, dtype=torch.float64).cuda()
        y = torch.randn(n, generator=g, dtype=torch.float64).cuda()
        want_f = k_xz @ w[:, :j]
        cost = nat.PlsCost()
        cost.cost_id, cost.link_id, cost.closed_form = nat.COST_STUDENT_T, nat.LINK_IDENTITY, 1
        cost.degrees_of_freedom, cost.scale, cost.link_jitter, cost.probit_divisor = 4.0, 0.7, 1e-10, 2.0 ** 0.5
        e = want_f - y[:, None]
        want_dc = 5.0 * e / (4.0 * 0.49 + e * e)
        want_c = (2.5 * torch.log(1.0 + e * e / (4.0 * 0.49))).sum(0)
        dcin = torch.randn(n, ld, generator=g, dtype=torch.float64).cuda()
        want_g = k_xz.T @ dcin[:, :j]
        wrap = torch.arange(256, device=xa.device) % n
        big = ops.gram_cache(ctx, nat.KERNEL_RBF, torch.cat([xa[wrap], xa, xa[wrap]]), za, d)
        big[:, m:] = 0.37

        def err(got, want):
            return (got - want).abs().max().item() / max(1.0, want.abs().max().item())

        for gram in (None, big[256:]):
            tr = ops.forward_tile_rows(ctx, j)
            tiles = (n + tr - 1) // tr
            f = torch.full((n, ld), 3.0, dtype=torch.float64).cuda()
            ops.forward(ctx, nat.KERNEL_RBF, xa, za, d, w, j, nat.EPI_PREDICTION, f, gram=gram)
            dc = torch.full((n, ld), 3.0, dtype=torch.float64).cuda()
            ops.forward(ctx, nat.KERNEL_RBF, xa, za, d, w, j, nat.EPI_COST_DERIVATIVE, dc, cost=cost, y=y, gram=gram)
            part = torch.zeros(tiles, ld, dtype=torch.float64).cuda()
            ops.forward(ctx, nat.KERNEL_RBF, xa, za, d, w, j, nat.EPI_COST, part, cost=cost, y=y, gram=gram)
            dc2 = torch.full((n, ld), 4.0, dtype=torch.float64).cuda()
            part2 = torch.zeros(tiles, ld, dtype=torch.float64).cuda()
            ops.forward_step(ctx, nat.KERNEL_RBF, xa, za, d, w, j, cost, y, dc2, part2, gram=gram)
            splits = ops.backward_splits(ctx, n, m, j)
            gp = torch.full((splits, m, ld), -2.0, dtype=torch.float64).cuda()
            ops.backward(ctx, nat.KERNEL_RBF, za, xa, d, dcin, j, gp, splits, accumulate=False, gram=gram)
            out = torch.empty(m, ld, dtype=torch.float64).cuda()
            ops.reduce_splits(ctx, gp, j, out)
            errs = [err(f[:, :j], want_f), err(dc[:, :j], want_dc), err(part[:, :j].sum(0), want_c), err(dc2[:, :j], want_dc),
                    err(part2[:, :j].sum(0), want_c), err(out[:, :j], want_g)]
            untouched = bool((f[:, j:] == 3.0).all() and (dc[:, j:] == 3.0).all() and (dc2[:, j:] == 4.0).all() and (gp[:, :, j:] == -2.0).all())
            worst = max(worst, max(errs))
            if max(errs) > 1e-11 or not untouched:
                print("FAIL", dict(n=n, m=m, d=d, j=j, ld=ld, rt=rt, cached=gram is not None), errs, untouched)
                sys.exit(1)
    ctx.lib.pls_set_tile_shape(ctx.handle, 0)
    ctx.lib.pls_set_tile_sets(ctx.handle, 0)
    print(f"sweep ok: {args.count} shapes x 2 Gram sources, worst scaled error {worst:.2e}")


if __name__ == "__main__":
    main()
