import torch


class MultivariateNormal(torch.distributions.MultivariateNormal):
    def __init__(self, mean, covariance_matrix, **kwargs):
        super().__init__(loc=mean, covariance_matrix=covariance_matrix, validate_args=False)


class base_distributions:
    StudentT = torch.distributions.StudentT


class Distribution:
    """annotation-only placeholder (src/gaussian_process/svgp.py:45)"""
