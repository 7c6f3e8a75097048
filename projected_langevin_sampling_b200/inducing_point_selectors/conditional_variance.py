"""ConditionalVariance inducing-point selector on the GPU
(reference: src/inducing_point_selectors/conditional_variance.py:11-120).

The greedy pivoted Cholesky runs in the CUDA library (csrc/pls_selector.cu): per iteration one memory-bound kernel
(column Gram + rank-1 update + clip + masked arg-max partials) and one tiny pivot kernel, with no host round trip.
The host keeps exactly what the reference does on the host: the numpy permutation from the GLOBAL generator
(conditional_variance.py:60), the final index mapping (:117-120) and -- only when the largest conditional variance is attained
by SEVERAL points to the last bit -- the reference's own tie-break, `reversed(np.argsort(d))` (:105-109, `ops.numpy_tie_rule`):
numpy's default argsort is unstable, so that choice is a property of numpy's sort routine and cannot be restated, only called.
`tie_rule="stable"` keeps every decision on the device (highest permuted index among the tied points)."""
from typing import Optional, Tuple

import numpy as np
import torch

from .. import _native as nat
from .. import ops
from ..kernels import kernel_spec
from .base import InducingPointSelector


class ConditionalVarianceInducingPointSelector(InducingPointSelector):
    def __init__(self, threshold: Optional[float] = 0.0, tie_rule="numpy"):
        self.threshold = threshold
        self.tie_rule = tie_rule  # "numpy" (the reference's choice on exact ties), "stable", or a callable (d, chosen) -> index
        self.last_run_info: dict = {}  # min_top2_rel_gap, tied_picks, host_tie_calls of the last compute_induce_data call

    def compute_induce_data(self, x: torch.Tensor, m: int, kernel, jitter: float = 1e-12) -> Tuple[torch.Tensor, torch.Tensor]:
        assert m > 1, "Must have at least 2 inducing points"
        if self.threshold is None:  # the reference compares `sum < None` and raises (conditional_variance.py:111)
            raise TypeError("'<' not supported between instances of 'float' and 'NoneType'")
        number_of_training_points = x.shape[0]
        perm = np.random.permutation(number_of_training_points)  # tie-breaking permutation, numpy GLOBAL generator (:60)
        ctx = nat.context()
        dev = torch.device("cuda", ctx.device_index)
        x_dev = ops.as_device_f64(x if x.dim() > 1 else x.unsqueeze(-1), dev)
        xp = x_dev[torch.from_numpy(perm).to(dev)]
        d = xp.shape[1]
        spec = kernel_spec(kernel, d)
        # centring only conditions the exponent arithmetic; the host mean of the UNPERMUTED x is used so that the sharded
        # variant below (where no rank holds all the permuted points) forms bit-identical augmented points
        centre = (x if x.dim() > 1 else x.unsqueeze(-1)).detach().double().cpu().mean(dim=0).tolist() if spec.kernel_id == nat.KERNEL_RBF else [0.0] * d
        # both sides of k(x, x_j) come from this one set: half of log(outputscale) on each side
        xa = ops.prepare_points(ctx, spec.kernel_id, xp, spec.inv_lengthscale, centre, 0.5 * spec.log_outputscale)
        self.last_run_info = {}
        local, n_selected = ops.cv_select(ctx, spec.kernel_id, xa, d, spec.outputscale, m, jitter, self.threshold,
                                          tie_rule=self.tie_rule, info=self.last_run_info)
        if n_selected < m:
            print("ConditionalVariance: Terminating selection of inducing points early.")
            # the reference indexes x with the untouched sentinel N here and raises (:63,117-118)
            raise IndexError(f"index {number_of_training_points} is out of bounds for dimension 0 with size {number_of_training_points}")
        local_cpu = local.cpu()
        indices = perm[local_cpu.numpy()]
        induce_data = x[torch.from_numpy(indices)] if not x.is_cuda else x[torch.from_numpy(indices).to(x.device)]
        return induce_data, torch.from_numpy(indices)

    def compute_induce_data_sharded(self, x: torch.Tensor, m: int, kernel, jitter: float = 1e-12, group=None,
                                    rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """The same selection with the N-sized state (C: (m-1) x N, d: N) sharded by rows over the ranks of `group` (one
        process per GPU): for N too large for one GPU (C is 655 GB at N = 20M, m = 4096).  Every rank passes the SAME x and
        has the SAME numpy seed (the permutation of conditional_variance.py:60 is drawn on every rank); rank r keeps rows
        shard_range(N, r, world) of the permuted set.  One all-gather of a (m + SP + 4)-double record per rank and pivot is
        the only communication.  Returns the reference's (x_induce, indices) on every rank, identical to the unsharded call."""
        import torch.distributed as dist

        from ..distributed import shard_range

        assert m > 1, "Must have at least 2 inducing points"
        if self.threshold is None:
            raise TypeError("'<' not supported between instances of 'float' and 'NoneType'")
        rank = dist.get_rank(group) if rank is None else rank
        world = dist.get_world_size(group) if world is None else world
        n = x.shape[0]
        perm = np.random.permutation(n)
        ctx = nat.context()
        dev = torch.device("cuda", ctx.device_index)
        x2 = x if x.dim() > 1 else x.unsqueeze(-1)
        d = x2.shape[1]
        spec = kernel_spec(kernel, d)
        r0, r1 = shard_range(n, rank, world)
        # the centre must be the same on every rank: the mean of the whole permuted set (= the mean of x)
        centre = x2.detach().double().cpu().mean(dim=0).tolist() if spec.kernel_id == nat.KERNEL_RBF else [0.0] * d
        xp_local = ops.as_device_f64(x2[torch.from_numpy(perm[r0:r1])], dev)
        xa = ops.prepare_points(ctx, spec.kernel_id, xp_local, spec.inv_lengthscale, centre, 0.5 * spec.log_outputscale)
        state = ops.ShardedSelectorState(ctx, spec.kernel_id, xa, r0, n, d, spec.outputscale, m, jitter, self.threshold,
                                         tie_rule=self.tie_rule)

        def gather(records):
            out = torch.empty((world * state.record,), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(out, records[0], group=group)
            return out

        def gather_d(slices):
            # the whole d on every rank's host (ties only): shards differ by at most one row, so pad to the longest
            longest = -(-n // world)
            padded = torch.zeros((longest,), dtype=torch.float64, device=dev)
            padded[: slices[0].numel()] = slices[0]
            out = torch.empty((world * longest,), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(out, padded, group=group)
            out = out.view(world, longest).cpu().numpy()
            return np.concatenate([out[r, : shard_range(n, r, world)[1] - shard_range(n, r, world)[0]] for r in range(world)])

        local, n_selected = ops.cv_select_sharded([state], gather, gather_d)
        if n_selected < m:
            print("ConditionalVariance: Terminating selection of inducing points early.")
            raise IndexError(f"index {n} is out of bounds for dimension 0 with size {n}")
        indices = perm[local.cpu().numpy()]
        induce_data = x[torch.from_numpy(indices)] if not x.is_cuda else x[torch.from_numpy(indices).to(x.device)]
        return induce_data, torch.from_numpy(indices)
