"""Stand-in for gpytorch.likelihoods.GaussianLikelihood (homoskedastic noise, batch_shape=[L] -> noise [L,1])."""
import torch

from .constraints import GreaterThan


class _HomoskedasticNoise(torch.nn.Module):
    def __init__(self, noise_constraint=None, batch_shape=torch.Size([])):
        super().__init__()
        self.register_parameter("raw_noise", torch.nn.Parameter(torch.zeros(*batch_shape, 1)))
        self.raw_noise_constraint = noise_constraint or GreaterThan(1e-4)

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        value = torch.as_tensor(value, dtype=self.raw_noise.dtype).expand(self.raw_noise.shape)
        with torch.no_grad():
            self.raw_noise.copy_(self.raw_noise_constraint.inverse_transform(value))


class GaussianLikelihood(torch.nn.Module):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size([]), **kwargs):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise(noise_constraint, torch.Size(batch_shape))

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    @property
    def raw_noise(self):
        return self.noise_covar.raw_noise
