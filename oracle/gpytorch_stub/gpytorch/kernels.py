import torch

from oracle.pls_oracle import RBFScaleKernel


LAZY_ABOVE = 50_000_000  # entries above which kernel(x1, x2) is returned unevaluated, as gpytorch always does


class _LazyGram:
    """What gpytorch returns from kernel(x1, x2): an unevaluated Gram.  The selector asks a 1M x 1M one only for its diagonal
    (conditional_variance.py:66-71: `kernel(x, x)`, `.ndim`, `.cpu().diagonal().detach().numpy()`), which gpytorch serves with
    the kernel's diag=True evaluation; anything else densifies."""

    def __init__(self, kernel, x1, x2, params):
        self.kernel, self.x1, self.x2, self.params = kernel, x1, x2, params
        self.ndim = 2
        self.shape = (x1.shape[0], x2.shape[0])

    def cpu(self):
        return self

    def diagonal(self):
        return self.kernel.forward(self.x1, self.x2, diag=True, **self.params)

    def __getattr__(self, name):
        return getattr(self.kernel.forward(self.x1, self.x2, **self.params), name)


class Kernel(torch.nn.Module):
    has_lengthscale = False

    def __init__(self, ard_num_dims=None, **kwargs):
        super().__init__()
        self.ard_num_dims = ard_num_dims
        if self.has_lengthscale:
            self.lengthscale = torch.ones(1, 1 if ard_num_dims is None else ard_num_dims, dtype=torch.get_default_dtype())

    def forward(self, x1, x2, diag=False, **params):
        raise NotImplementedError

    def __call__(self, x1, x2=None, diag=False, **params):
        x2 = x1 if x2 is None else x2
        if x1.ndim == 1:
            x1 = x1.unsqueeze(1)
        if x2.ndim == 1:
            x2 = x2.unsqueeze(1)
        if not diag and x1.shape[0] * x2.shape[0] > LAZY_ABOVE:
            return _LazyGram(self, x1, x2, params)
        return self.forward(x1, x2, diag=diag, **params)

    def cuda(self):
        return self


class RBFKernel(Kernel):
    has_lengthscale = True

    def forward(self, x1, x2, diag=False, **params):
        return RBFScaleKernel(lengthscale=self.lengthscale, outputscale=1.0)(x1, x2, diag=diag)


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, **kwargs):
        super().__init__(**kwargs)
        self.base_kernel = base_kernel
        self.outputscale = 1.0

    def forward(self, x1, x2, diag=False, **params):
        return self.base_kernel.forward(x1, x2, diag=diag, **params).mul(float(self.outputscale))
