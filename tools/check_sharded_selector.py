"""Multi-GPU check of the row-sharded ConditionalVariance selector: run under torchrun, one rank per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
        tools/check_sharded_selector.py --points 400000 --dim 8 --inducing 256

Every rank runs the unsharded selector on its own GPU and the sharded one over NCCL; the indices must be identical.
"""
import argparse
import json
import math
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import projected_langevin_sampling_b200 as pkg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", dest="n", type=int, default=400_000)
    ap.add_argument("--dim", dest="d", type=int, default=8)
    ap.add_argument("--inducing", dest="m", type=int, default=256)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = torch.Generator().manual_seed(0)
    x = torch.randn(args.n, args.d, generator=g, dtype=torch.float64)
    ls = torch.tensor([math.sqrt(args.d) * (0.75 + 0.5 * k / max(args.d - 1, 1)) for k in range(args.d)], dtype=torch.float64)
    kernel = pkg.ScaleKernel(pkg.RBFKernel(ard_num_dims=args.d, lengthscale=ls), outputscale=1.0)
    sel = pkg.ConditionalVarianceInducingPointSelector()
    pkg.set_seed(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    z_one, idx_one = sel(x=x, m=args.m, kernel=kernel)
    torch.cuda.synchronize()
    t_one = time.perf_counter() - t0
    pkg.set_seed(3)
    dist.barrier()
    t0 = time.perf_counter()
    z_sh, idx_sh = sel.compute_induce_data_sharded(x, args.m, kernel)
    torch.cuda.synchronize()
    dist.barrier()
    t_sh = time.perf_counter() - t0
    same = torch.tensor([int(torch.equal(idx_one, idx_sh) and torch.equal(z_one, z_sh))], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "n": args.n, "d": args.d, "m": args.m, "identical_indices": bool(same.item()),
                          "seconds_one_gpu": round(t_one, 4), "seconds_sharded": round(t_sh, 4), "first_indices": idx_sh[:6].tolist()}))
        if not same.item():
            sys.exit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
