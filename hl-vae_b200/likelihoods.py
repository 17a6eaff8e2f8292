"""GaussianLikelihood with the surface HLVAE_main.py:208-213 and elbo_functions.py:151,250 use:
`noise_covar.noise` -> [L, 1], `.noise = v` setter, `.raw_noise`.  noise = softplus(raw) + lower
bound, as gpytorch's GaussianLikelihood with a GreaterThan constraint.  Any object exposing
`noise_covar.noise` (e.g. the real gpytorch likelihood) is accepted by the KL functions."""
from __future__ import annotations

import torch
from torch import nn

from .kernels import _Positive


class GreaterThan(_Positive):
    def __init__(self, lower_bound):
        super().__init__(lower_bound)


class _HomoskedasticNoise(nn.Module):
    def __init__(self, batch_shape, constraint):
        super().__init__()
        self.register_parameter("raw_noise", nn.Parameter(torch.zeros(*batch_shape, 1)))
        self.raw_noise_constraint = constraint

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        value = torch.as_tensor(value).to(self.raw_noise)
        with torch.no_grad():
            self.raw_noise.copy_(self.raw_noise_constraint.inverse_transform(value).expand_as(self.raw_noise))


class GaussianLikelihood(nn.Module):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size([]), **kwargs):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise(torch.Size(batch_shape), noise_constraint or GreaterThan(1e-4))

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    @property
    def raw_noise(self):
        return self.noise_covar.raw_noise
