"""ConditionalVariance inducing-point selector on the GPU
(reference: src/inducing_point_selectors/conditional_variance.py:11-120).

The greedy pivoted Cholesky runs in the CUDA library (csrc/pls_selector.cu): per iteration one memory-bound kernel
(column Gram + rank-1 update + clip + masked arg-max partials) and one tiny pivot kernel, with no host round trip.
The host keeps exactly what the reference does on the host: the numpy permutation from the GLOBAL generator
(conditional_variance.py:60) and the final index mapping (:117-120)."""
from typing import Optional, Tuple

import numpy as np
import torch

from .. import _native as nat
from .. import ops
from ..kernels import kernel_spec
from .base import InducingPointSelector


class ConditionalVarianceInducingPointSelector(InducingPointSelector):
    def __init__(self, threshold: Optional[float] = 0.0):
        self.threshold = threshold

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
        centre = xp.mean(dim=0).tolist() if spec.kernel_id == nat.KERNEL_RBF else [0.0] * d
        # both sides of k(x, x_j) come from this one set: half of log(outputscale) on each side
        xa = ops.prepare_points(ctx, spec.kernel_id, xp, spec.inv_lengthscale, centre, 0.5 * spec.log_outputscale)
        local, n_selected = ops.cv_select(ctx, spec.kernel_id, xa, d, spec.outputscale, m, jitter, self.threshold)
        if n_selected < m:
            print("ConditionalVariance: Terminating selection of inducing points early.")
            # the reference indexes x with the untouched sentinel N here and raises (:63,117-118)
            raise IndexError(f"index {number_of_training_points} is out of bounds for dimension 0 with size {number_of_training_points}")
        local_cpu = local.cpu()
        indices = perm[local_cpu.numpy()]
        induce_data = x[torch.from_numpy(indices)] if not x.is_cuda else x[torch.from_numpy(indices).to(x.device)]
        return induce_data, torch.from_numpy(indices)
