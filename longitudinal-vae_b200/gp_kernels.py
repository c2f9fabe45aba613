"""Kernel modules with GPyTorch's interface and state_dict keys, evaluated by the CUDA dense-kernel op.

The reference builds its kernels from `gpytorch.kernels.{Kernel, RBFKernel, ScaleKernel, ProductKernel, AdditiveKernel}`
(kernel_spec.py:2-3, kernel_gen.py:3).  GPyTorch is a third-party dependency that this package does not require: the
classes below keep the attribute names the reference relies on (`kernels`, `base_kernel`, `outputscale`, `lengthscale`,
`active_dims`, `raw_*` parameters with softplus constraints, `k1 * k2`, `k0 + k1`, `module(x1, x2).evaluate()`), but hold
no dense torch arithmetic — `.evaluate()` flattens the tree (spec.py) and runs lvae_kernel_dense_f64 on the GPU
(lvae_kernel_dense_bwd_f64 under autograd).
"""
import torch
from torch.nn import ModuleList

from .constraints import Positive
from .constraints import GreaterThan
from .spec import FlatComponent, Raw, build_structure, flatten, latent_count


def _lazy(raw, constraint, evaluated):
    """A Raw reference when the constraint is the plain softplus + lower bound (packed evaluation, spec.build_structure),
    otherwise the evaluated tensor."""
    if type(constraint) in (GreaterThan, Positive):
        return Raw(raw, "softplus", constraint.lower_float())
    return evaluated()


class LazyKernelTensor:
    """Result of `kernel(x1, x2)`; `.evaluate()` / `.to_dense()` gives the dense matrix (GPyTorch's lazy contract)."""

    def __init__(self, kernel, x1, x2):
        self.kernel, self.x1, self.x2 = kernel, x1, x2

    def evaluate(self):
        return evaluate_dense(self.kernel, self.x1, self.x2)

    to_dense = evaluate


def evaluate_dense(kernel, x1, x2):
    """Dense kernel matrix with the reference's broadcasting: x [n,Q] | [L,n,Q] | [P,L,n,Q] against [L,1,1] parameters
    gives [L,n1,n2] | [P,L,n1,n2]; un-batched kernels (no latent batch) give [n1,n2].
    Differentiable w.r.t. the hyper-parameters (diff_ops.KernelDense), constant in the covariates."""
    from .diff_ops import KernelDense
    comps = flatten(kernel)
    L = latent_count(comps, default=1)
    batched = any(t is not None and (torch.is_tensor(t) or isinstance(t, Raw)) and t.numel() > 1
                  for c in comps for t in [c.outputscale] + [f[2] for f in c.factors]) or \
        len(getattr(kernel, "batch_shape", ())) > 0
    dev = x1.device
    structure, ls, os_ = build_structure(comps, [], L, device=dev)
    lead = max(x1.dim(), x2.dim())
    if lead == 4:                                   # [P,L,n,Q] stacks of elbo_functions.py:168-174
        P = x1.shape[0] if x1.dim() == 4 else x2.shape[0]
        f1 = x1.expand(P, L, *x1.shape[-2:]).reshape(P * L, *x1.shape[-2:])
        f2 = x2.expand(P, L, *x2.shape[-2:]).reshape(P * L, *x2.shape[-2:])
        out = KernelDense.apply(structure, "all", f1, f2, ls, os_, None)
        return out.view(P, L, out.shape[-2], out.shape[-1])
    a = x1 if x1.dim() < 3 or x1.shape[0] == L else x1.expand(L, *x1.shape[-2:])
    b = x2 if x2.dim() < 3 or x2.shape[0] == L else x2.expand(L, *x2.shape[-2:])
    out = KernelDense.apply(structure, "all", a, b, ls, os_, None)
    if lead == 2 and not batched:
        return out[0]
    return out


class Kernel(torch.nn.Module):
    has_lengthscale = False

    def __init__(self, has_lengthscale=False, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None,
                 lengthscale_constraint=None, eps=1e-6, **kwargs):
        super().__init__()
        self._batch_shape = torch.Size(batch_shape)
        self._dim_int = None                      # host copy of the covariate column (no device sync on the hot path)
        if active_dims is not None:
            self._dim_int = int(torch.as_tensor(active_dims).reshape(-1)[0])
            if not torch.is_tensor(active_dims):
                active_dims = torch.tensor(active_dims, dtype=torch.long)
        self.register_buffer("active_dims", active_dims)
        self.ard_num_dims = ard_num_dims
        self.eps = eps
        if has_lengthscale or type(self).has_lengthscale:
            self.register_parameter("raw_lengthscale", torch.nn.Parameter(torch.zeros(*self._batch_shape, 1, 1)))
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
            value = torch.as_tensor(value, dtype=raw.dtype, device=raw.device).expand(raw.shape)
            with torch.no_grad():
                raw.copy_(cons.inverse_transform(value))
        return self

    def _dim(self):
        if self._dim_int is None:
            if self.active_dims is None:
                raise ValueError("lvae_b200: leaf kernels need active_dims (a covariate column)")
            self._dim_int = int(self.active_dims.reshape(-1)[0])
        return self._dim_int

    def forward(self, x1, x2, **params):
        return evaluate_dense(self, x1, x2)

    def __call__(self, x1, x2=None, **params):
        return LazyKernelTensor(self, x1, x1 if x2 is None else x2)

    def __add__(self, other):
        ks = (list(self.kernels) if isinstance(self, AdditiveKernel) else [self]) + \
             (list(other.kernels) if isinstance(other, AdditiveKernel) else [other])
        return AdditiveKernel(*ks)

    def __mul__(self, other):
        ks = (list(self.kernels) if isinstance(self, ProductKernel) else [self]) + \
             (list(other.kernels) if isinstance(other, ProductKernel) else [other])
        return ProductKernel(*ks)


class RBFKernel(Kernel):
    """exp(-(x1-x2)^2 / (2 l^2)) on one covariate column; l = softplus(raw_lengthscale) per latent."""
    has_lengthscale = True

    def _flat_components(self):
        return [FlatComponent(None, [('rbf', self._dim(), _lazy(self.raw_lengthscale, self.raw_lengthscale_constraint,
                                                               lambda: self.lengthscale.reshape(-1)))])]


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, outputscale_constraint=None, **kwargs):
        if getattr(base_kernel, "active_dims", None) is not None:
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

    def _flat_components(self):
        scale = FlatComponent(_lazy(self.raw_outputscale, self.raw_outputscale_constraint,
                                    lambda: self.outputscale.reshape(-1)), [])
        return [scale.times(c) for c in flatten(self.base_kernel)]


class ProductKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = ModuleList(kernels)

    def _flat_components(self):
        out = None
        for k in self.kernels:
            cs = flatten(k)
            out = cs if out is None else [a.times(b) for a in out for b in cs]
        return out or []


class AdditiveKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = ModuleList(kernels)

    def _flat_components(self):
        return [c for k in self.kernels for c in flatten(k)]
