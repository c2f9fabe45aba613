"""Drop-in for the reference's kernel_spec.py: leaf kernels of the additive GP prior."""
from .gp_kernels import Kernel, RBFKernel
from .spec import FlatComponent


class BinKernel(Kernel):
    """Binary kernel 1[x1 + x2 == 2] on one covariate column (kernel_spec.py:9-23; `value` is stored, unused in forward)."""

    def __init__(self, value, **kwargs):
        super().__init__(has_lengthscale=False, **kwargs)
        self.value = value

    def _flat_components(self):
        return [FlatComponent(None, [('bin', self._dim(), None)])]


class CatKernel(Kernel):
    """Categorical kernel 1[x1 - x2 == 0] on one covariate column (kernel_spec.py:26-32)."""

    def _flat_components(self):
        return [FlatComponent(None, [('cat', self._dim(), None)])]


class CatKernelMod(Kernel):
    """Categorical kernel with -1/(num-1) off the diagonal (kernel_spec.py:35-55).  No generator of the reference uses
    it; it equals (1 + 1/(num-1)) * cat - 1/(num-1) and is evaluated from the CUDA categorical mask."""

    def __init__(self, num, **kwargs):
        super().__init__(has_lengthscale=False, **kwargs)
        self.num = num

    def _flat_components(self):
        raise TypeError("lvae_b200: CatKernelMod cannot be part of a fused additive kernel (unused by the reference)")

    def forward(self, x1, x2, **params):
        cat = CatKernel(active_dims=self.active_dims)
        same = cat(x1.reshape(-1, 1) if self.active_dims is None else x1, x2.reshape(-1, 1) if self.active_dims is None else x2).evaluate()
        off = -1.0 / (self.num - 1)
        return same * (1.0 - off) + off

    def __call__(self, x1, x2=None, **params):
        x2 = x1 if x2 is None else x2
        outer = self

        class _Lazy:
            def evaluate(self_inner):
                return outer.forward(x1, x2)
            to_dense = evaluate
        return _Lazy()


def RbfKernel(active_dims, batch_shape=None):
    """Squared-exponential kernel on one covariate column, lengthscale 2.5 per latent to start with (kernel_spec.py:58-69)."""
    shape = {} if batch_shape is None else {"batch_shape": batch_shape}
    return RBFKernel(active_dims=active_dims, **shape).initialize(lengthscale=2.5)
