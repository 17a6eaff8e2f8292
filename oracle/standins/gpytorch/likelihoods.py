import torch
from .constraints import GreaterThan


class HomoskedasticNoise(torch.nn.Module):
    def __init__(self, batch_shape, noise_constraint):
        super().__init__()
        self.register_parameter("raw_noise", torch.nn.Parameter(torch.zeros(*batch_shape, 1)))
        self.raw_noise_constraint = noise_constraint

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        value = torch.as_tensor(value).to(self.raw_noise)
        with torch.no_grad():
            self.raw_noise.copy_(self.raw_noise_constraint.inverse_transform(value).expand_as(self.raw_noise))


class GaussianLikelihood(torch.nn.Module):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size([]), **kwargs):
        super().__init__()
        self.noise_covar = HomoskedasticNoise(batch_shape, noise_constraint or GreaterThan(1e-4))

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    @property
    def raw_noise(self):
        return self.noise_covar.raw_noise
