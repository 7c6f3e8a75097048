"""Multi-GPU check of the (row shards x particle shards) grid: run under torchrun with one rank per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        tools/check_row_sharding.py --grid 2x1

Every rank holds a row slice of (X, y) and a particle slice of P; per step the (M x J_local) gradient is all-reduced over
the ranks that share a particle slice (NCCL), the Philox noise is keyed on the global particle index.  Every rank also
runs the whole problem on its own GPU and compares its slice: the sharded run must reproduce the single-GPU run
(gradient summation order differs, so agreement is to round-off, not bitwise).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import projected_langevin_sampling_b200 as pkg  # noqa: E402
from projected_langevin_sampling_b200.distributed import GridPlacement, gradient_allreduce, make_row_group  # noqa: E402
from projected_langevin_sampling_b200.projected_langevin_sampling import costs, link_functions as lf  # noqa: E402
from projected_langevin_sampling_b200.trainers import train_pls  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=str, default="2x1", help="row shards x particle shards")
    ap.add_argument("--points", dest="n", type=int, default=30000)
    ap.add_argument("--inducing", dest="m", type=int, default=128)
    ap.add_argument("--dim", dest="d", type=int, default=4)
    ap.add_argument("--particles", dest="j", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_groups, j_groups = (int(v) for v in args.grid.split("x"))
    place = GridPlacement(rank=rank, world=world, n_groups=n_groups, j_groups=j_groups)
    group = make_row_group(place)

    g = torch.Generator().manual_seed(0)
    n, m, d, j = args.n, args.m, args.d, args.j
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    z = x[:m].clone()
    ls = torch.tensor([1.5 + 0.2 * k for k in range(d)], dtype=torch.float64)
    kernel = pkg.ScaleKernel(pkg.RBFKernel(ard_num_dims=d, lengthscale=ls), outputscale=1.2)
    p0 = torch.randn(m, j, generator=g, dtype=torch.float64)  # rows trimmed to M_k below
    eta, seed = 1e-4, 99

    def build(xr, yr, reduce_hook):
        basis = pkg.OrthonormalBasis(pkg.PLSKernel(kernel, z), z, xr, eigenvalue_threshold=1e-9, gradient_reduce=reduce_hook, verbose=False)
        return pkg.PLS(basis, costs.GaussianCost(0.05, yr, lf.IdentityLinkFunction()))

    # the whole problem on this GPU
    full = build(x, y, None)
    m_k = full.basis.approximation_dimension
    p_full = p0[:m_k].cuda()
    want_energy = []
    for s in range(args.steps):
        full.step_(p_full, eta, philox=(seed, s, 0))
        want_energy.append(full.calculate_energy_potential(p_full))

    # this rank's shard of the grid
    r0, r1 = place.rows(n)
    j0, j1 = place.particles(j)
    shard = build(x[r0:r1], y[r0:r1], gradient_allreduce(group))
    p = p0[:m_k, j0:j1].contiguous().cuda()
    for s in range(args.steps):
        shard.step_(p, eta, philox=(seed, s, j0))
    err = ((p - p_full[:, j0:j1]).abs().max() / p_full.abs().max()).item()

    # the fused training loop under row sharding: per-particle energies are summed over the row group inside train_pls;
    # the mean over ALL particles needs the particle shards' means combined (equal shard sizes here)
    p_t = p0[:m_k, j0:j1].contiguous().cuda()
    _, energies = train_pls(shard, p_t, args.steps, eta, early_stopper_patience=1e9, philox_seed=seed, j_global_offset=j0)
    e = torch.tensor(energies, dtype=torch.float64, device="cuda") * (j1 - j0)
    # sum the per-shard particle sums over ONE representative of every particle shard (the ranks of row shard 0)
    contrib = e if place.n_index == 0 else torch.zeros_like(e)
    dist.all_reduce(contrib)
    got_energy = (contrib / j).tolist()
    e_err = max(abs(a - b) / abs(b) for a, b in zip(got_energy, want_energy))
    errs = torch.tensor([err, e_err, (p_t - p_full[:, j0:j1]).abs().max().item() / p_full.abs().max().item()], dtype=torch.float64, device="cuda")
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if rank == 0:
        ok = bool((errs < 1e-11).all())
        print(json.dumps({"grid": args.grid, "world": world, "n": n, "m": m, "m_k": m_k, "j": j, "steps": args.steps,
                          "max_rel_err_particles": errs[0].item(), "max_rel_err_energy": errs[1].item(),
                          "max_rel_err_particles_train_pls": errs[2].item(), "ok": ok}))
        if not ok:
            sys.exit(1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
