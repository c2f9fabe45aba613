"""Stand-in for gpytorch.kernels: dense semantics of Kernel / RBFKernel / ScaleKernel / ProductKernel / AdditiveKernel."""
import torch
from torch.nn import ModuleList

from .constraints import Positive


class _Dense:
    """What Kernel.__call__ returns in GPyTorch is lazy; `.evaluate()` (later `.to_dense()`) densifies it."""

    def __init__(self, fn):
        self._fn = fn

    def evaluate(self):
        return self._fn()

    to_dense = evaluate


def _densify(obj):
    return obj.evaluate() if isinstance(obj, _Dense) else obj


class Kernel(torch.nn.Module):
    has_lengthscale = False

    def __init__(self, has_lengthscale=False, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None,
                 lengthscale_constraint=None, eps=1e-6, **kwargs):
        super().__init__()
        self._batch_shape = torch.Size(batch_shape)
        if active_dims is not None and not torch.is_tensor(active_dims):
            active_dims = torch.tensor(active_dims, dtype=torch.long)
        self.register_buffer("active_dims", active_dims)
        self.ard_num_dims = ard_num_dims
        self.eps = eps
        if has_lengthscale or type(self).has_lengthscale:
            n = 1 if ard_num_dims is None else ard_num_dims
            self.register_parameter("raw_lengthscale", torch.nn.Parameter(torch.zeros(*self._batch_shape, 1, n)))
            self.raw_lengthscale_constraint = lengthscale_constraint or Positive()

    @property
    def batch_shape(self):
        return self._batch_shape

    @property
    def lengthscale(self):
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, value):
        self.initialize(lengthscale=value)

    def initialize(self, **kwargs):
        for name, value in kwargs.items():
            raw = getattr(self, "raw_" + name)
            cons = getattr(self, "raw_" + name + "_constraint")
            value = torch.as_tensor(value, dtype=raw.dtype).expand(raw.shape)
            with torch.no_grad():
                raw.copy_(cons.inverse_transform(value))
        return self

    def forward(self, x1, x2, **params):
        raise NotImplementedError

    def __call__(self, x1, x2=None, **params):
        x1_, x2_ = x1, (x1 if x2 is None else x2)
        if self.active_dims is not None:
            x1_ = x1_.index_select(-1, self.active_dims)
            x2_ = x2_.index_select(-1, self.active_dims)
        if x1_.ndimension() == 1:
            x1_ = x1_.unsqueeze(1)
        if x2_.ndimension() == 1:
            x2_ = x2_.unsqueeze(1)
        return _Dense(lambda: _densify(torch.nn.Module.__call__(self, x1_, x2_, **params)))

    def __add__(self, other):
        ks = (list(self.kernels) if isinstance(self, AdditiveKernel) else [self]) + \
             (list(other.kernels) if isinstance(other, AdditiveKernel) else [other])
        return AdditiveKernel(*ks)

    def __mul__(self, other):
        ks = (list(self.kernels) if isinstance(self, ProductKernel) else [self]) + \
             (list(other.kernels) if isinstance(other, ProductKernel) else [other])
        return ProductKernel(*ks)


def _sq_dist(x1, x2):
    """gpytorch.kernels.kernel.sq_dist: mean-centred ||a||^2 + ||b||^2 - 2ab^T expansion, clamped at 0."""
    adjustment = x1.mean(-2, keepdim=True)
    x1 = x1 - adjustment
    x2 = x2 - adjustment
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
    x1_ = torch.cat([-2.0 * x1, x1_norm, torch.ones_like(x1_norm)], dim=-1)
    x2_ = torch.cat([x2, torch.ones_like(x2_norm), x2_norm], dim=-1)
    return x1_.matmul(x2_.transpose(-2, -1)).clamp_min(0)


class RBFKernel(Kernel):
    has_lengthscale = True

    def forward(self, x1, x2, **params):
        x1_ = x1.div(self.lengthscale)
        x2_ = x2.div(self.lengthscale)
        return _sq_dist(x1_, x2_).div(-2).exp()


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, outputscale_constraint=None, **kwargs):
        if base_kernel.active_dims is not None:
            kwargs["active_dims"] = base_kernel.active_dims
        super().__init__(**kwargs)
        self.base_kernel = base_kernel
        shape = self.batch_shape
        init = torch.zeros(*shape) if len(shape) else torch.tensor(0.0)
        self.register_parameter("raw_outputscale", torch.nn.Parameter(init))
        self.raw_outputscale_constraint = outputscale_constraint or Positive()

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        self.initialize(outputscale=value)

    def forward(self, x1, x2, **params):
        orig = _densify(self.base_kernel.forward(x1, x2, **params))
        scales = self.outputscale
        return orig.mul(scales.view(*scales.shape, 1, 1))


class ProductKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = ModuleList(kernels)

    def forward(self, x1, x2, **params):
        res = _densify(self.kernels[0](x1, x2, **params))
        for kern in self.kernels[1:]:
            res = res * _densify(kern(x1, x2, **params))
        return res


class AdditiveKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = ModuleList(kernels)

    def forward(self, x1, x2, **params):
        res = 0
        for kern in self.kernels:
            res = res + _densify(kern(x1, x2, **params))
        return res
