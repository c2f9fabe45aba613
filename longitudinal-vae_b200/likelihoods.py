"""Gaussian likelihood container with GPyTorch's attribute names (LVAE.py:183-188): `noise_covar.noise` is [L,1] =
softplus(raw_noise) + lower bound; `likelihood.noise = 1` re-initialises raw_noise."""
import torch

from .constraints import GreaterThan


class HomoskedasticNoise(torch.nn.Module):
    def __init__(self, noise_constraint=None, batch_shape=torch.Size([])):
        super().__init__()
        self.register_parameter("raw_noise", torch.nn.Parameter(torch.zeros(*batch_shape, 1)))
        self.raw_noise_constraint = noise_constraint or GreaterThan(1e-4)

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        value = torch.as_tensor(value, dtype=self.raw_noise.dtype, device=self.raw_noise.device).expand(self.raw_noise.shape)
        with torch.no_grad():
            self.raw_noise.copy_(self.raw_noise_constraint.inverse_transform(value))


class GaussianLikelihood(torch.nn.Module):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size([]), **kwargs):
        super().__init__()
        self.noise_covar = HomoskedasticNoise(noise_constraint, torch.Size(batch_shape))

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    @property
    def raw_noise(self):
        return self.noise_covar.raw_noise
