"""Drop-in for the reference's GP_model.py (its gpytorch-free kernel/likelihood restatement, GP_model.py:7-236).

Same class names, constructor arguments, parameter names (`_log_noise`, `_log_lengthscale`, `_log_scale`) and buffers
(`min_log_*`); values are `exp(min_log + softplus(raw - min_log))` with `min_log = -16` and float32 initialisation as in
the reference (GP_model.py:16-18,65-67,97-99).  `forward(x1, x2)` returns a dense tensor computed by the CUDA
dense-kernel op; latent dimension first (`[L, n1, n2]`), as the reference's classes broadcast.
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

from .spec import FlatComponent, Raw, build_structure, flatten, latent_count


def _min_float(module, name):
    """A `min_log_*` buffer as a Python float, read once (no device sync on the hot path)."""
    key = "_f_" + name
    v = module.__dict__.get(key)
    if v is None:
        v = float(getattr(module, name))
        module.__dict__[key] = v
    return v


def _bounded(raw, min_log):
    return torch.exp(min_log + F.softplus(raw - min_log))


_FLOOR = -16.0


def _add_bounded(module, raw_name, floor_name, value, latent_dim, trainable=True):
    """Registers `raw_name` ([latent_dim] parameter, float32 like the reference's) and the `floor_name` buffer so that
    exp(floor + softplus(raw - floor)) starts at `value`."""
    floor = torch.full((1,), _FLOOR)
    start = torch.log(torch.as_tensor(value, dtype=torch.float32) - torch.exp(floor)).reshape(())
    setattr(module, raw_name, nn.Parameter(start.repeat(latent_dim), requires_grad=trainable))
    module.register_buffer(floor_name, floor)


def _bounded_property(raw_name, floor_name):
    """value = exp(floor + softplus(raw - floor)); assigning a value re-initialises the raw parameter."""
    def read(self):
        return _bounded(getattr(self, raw_name), getattr(self, floor_name))

    def write(self, value):
        with torch.no_grad():
            getattr(self, raw_name).copy_(torch.log(torch.as_tensor(value) - torch.exp(getattr(self, floor_name))))
    return property(read, write)


class Likelihoods(nn.Module):
    def __init__(self, latent_dim, noise, constrain=True):
        super().__init__()
        self.latent_dim = latent_dim
        _add_bounded(self, '_log_noise', 'min_log_noise', noise, latent_dim, trainable=constrain)

    noise = _bounded_property('_log_noise', 'min_log_noise')


class _DenseModule(nn.Module):
    def forward(self, x1, x2):
        comps = flatten(self)
        L = latent_count(comps, default=1)
        structure, ls, os_ = build_structure(comps, [], L, device=x1.device)
        a = x1 if x1.dim() < 3 or x1.shape[0] == L else x1.expand(L, *x1.shape[-2:])
        b = x2 if x2.dim() < 3 or x2.shape[0] == L else x2.expand(L, *x2.shape[-2:])
        from .diff_ops import KernelDense
        return KernelDense.apply(structure, "all", a, b, ls, os_, None)


class BinKernel(_DenseModule):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def _flat_components(self):
        return [FlatComponent(None, [('bin', self.dim, None)])]


class CatKernel(_DenseModule):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def _flat_components(self):
        return [FlatComponent(None, [('cat', self.dim, None)])]


class RbfKernel(_DenseModule):
    def __init__(self, dim, latent_dim=1, lengthscale=2.5):
        super().__init__()
        self.dim = dim
        self.latent_dim = latent_dim
        _add_bounded(self, '_log_lengthscale', 'min_log_lengthscale', lengthscale, latent_dim)

    lengthscale = _bounded_property('_log_lengthscale', 'min_log_lengthscale')

    def _flat_components(self):
        return [FlatComponent(None, [('rbf', self.dim, Raw(self._log_lengthscale, 'bounded', _min_float(self, 'min_log_lengthscale')))])]


class ScaleKernel(_DenseModule):
    def __init__(self, kernel, latent_dim=1, scale=math.log(2)):
        super().__init__()
        self.latent_dim = latent_dim
        self.kernel = kernel
        _add_bounded(self, '_log_scale', 'min_log_scale', scale, latent_dim)

    scale = _bounded_property('_log_scale', 'min_log_scale')

    def _flat_components(self):
        s = FlatComponent(Raw(self._log_scale, 'bounded', _min_float(self, 'min_log_scale')), [])
        return [s.times(c) for c in flatten(self.kernel)]


class AdditiveKernel(_DenseModule):
    def __init__(self, kernels):
        super().__init__()
        self.kernels = nn.ModuleList(kernels)

    def _flat_components(self):
        return [c for k in self.kernels for c in flatten(k)]


class ProductKernel(_DenseModule):
    def __init__(self, kernel1, kernel2):
        super().__init__()
        self.k1 = kernel1
        self.k2 = kernel2

    def _flat_components(self):
        return [a.times(b) for a in flatten(self.k1) for b in flatten(self.k2)]


def generate_kernel_batched(latent_dim, cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel,
                            covariate_missing_val, id_covariate):
    """(K0, K1) AdditiveKernels; same K0/K1 assignment and component order as kernel_gen (GP_model.py:146-236)."""
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    missing = [d['covariate'] for d in covariate_missing_val]

    def masked(kernel, covariate):
        if covariate in missing:
            return ProductKernel(kernel, BinKernel(covariate_missing_val[missing.index(covariate)]['mask']))
        return kernel

    k0, k1 = [], []
    for d in cat_kernel:
        (k1 if d == id_covariate else k0).append(ScaleKernel(masked(CatKernel(d), d), latent_dim))
    for d in sqexp_kernel:
        k0.append(ScaleKernel(masked(RbfKernel(d, latent_dim), d), latent_dim))
    for d in bin_kernel:
        k0.append(ScaleKernel(masked(BinKernel(d), d), latent_dim))
    for e in cat_int_kernel:
        prod = ProductKernel(masked(CatKernel(e['cat_covariate']), e['cat_covariate']),
                             masked(RbfKernel(e['cont_covariate'], latent_dim), e['cont_covariate']))
        (k1 if e['cat_covariate'] == id_covariate else k0).append(ScaleKernel(prod, latent_dim))
    for e in bin_int_kernel:
        prod = ProductKernel(masked(BinKernel(e['bin_covariate']), e['bin_covariate']),
                             masked(RbfKernel(e['cont_covariate'], latent_dim), e['cont_covariate']))
        k0.append(ScaleKernel(prod, latent_dim))
    return AdditiveKernel(k0).to(device), AdditiveKernel(k1).to(device)
