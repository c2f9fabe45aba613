"""Container with the state_dict key layout of the reference's GP_def.ExactGPModel (GP_def.py:8-21), which in the Hensman
path only exists to hold `likelihood` + `covar_module` for `gp_model.state_dict()` / `load_state_dict()` / `.train()`
(LVAE.py:196-197, 213-218, 353-360; training.py:199-204).  Keys: `likelihood.noise_covar.raw_noise`,
`covar_module.kernels.<i>.raw_outputscale`, `covar_module.kernels.<i>.base_kernel[.kernels.<j>].raw_lengthscale`, the
constraints' `lower_bound` / `upper_bound` buffers and the `active_dims` buffers — gpytorch's layout, so a `gp_model.pth`
written by either side loads on the other.  Exact-GP inference (`model(x)` -> MultivariateNormal, used by the reference
only outside the Hensman path) is out of scope: forward gives the zero mean and the dense prior covariance."""
import os

import torch


class ZeroMean(torch.nn.Module):
    def forward(self, x):
        return torch.zeros(x.shape[:-1], dtype=x.dtype, device=x.device)


class ExactGPModel(torch.nn.Module):
    def __init__(self, train_x, train_y, likelihood, covar_module):
        super().__init__()
        self.train_inputs = (train_x,) if train_x is not None else None
        self.train_targets = train_y
        self.likelihood = likelihood
        self.mean_module = ZeroMean()
        self.covar_module = covar_module

    def forward(self, x):
        return self.mean_module(x), self.covar_module(x, x).evaluate()


# file names of LVAE.py:353-360 (suffix "" at the end of training) and training.py:199-204 (suffix "_best")
def save_hensman_state(folder, gp_model, zt_list, m, H, suffix=""):
    """torch.save the four files the reference writes for the Hensman path."""
    torch.save(gp_model.state_dict(), os.path.join(folder, f"gp_model{suffix}.pth"))
    torch.save(zt_list, os.path.join(folder, f"zt_list{suffix}.pth"))
    torch.save(m, os.path.join(folder, f"m{suffix}.pth"))
    torch.save(H, os.path.join(folder, f"H{suffix}.pth"))


def load_hensman_state(folder, gp_model, device, suffix=""):
    """Counterpart of LVAE.py:213-232: loads the state_dict strictly into gp_model and returns (zt_list, m, H) on `device`."""
    dev = torch.device(device)
    gp_model.load_state_dict(torch.load(os.path.join(folder, f"gp_model{suffix}.pth"), map_location=dev))
    zt_list = torch.load(os.path.join(folder, f"zt_list{suffix}.pth"), map_location=dev)
    m = torch.load(os.path.join(folder, f"m{suffix}.pth"), map_location=dev).detach()
    H = torch.load(os.path.join(folder, f"H{suffix}.pth"), map_location=dev).detach()
    return zt_list, m, H
