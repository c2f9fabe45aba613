"""TEST INFRASTRUCTURE ONLY — minimal stand-in for the slice of GPyTorch (>=1.3) the reference touches.

GPyTorch is a third-party dependency of the reference (README.MD:25, "gpytorch >= 1.3", unpinned) and is neither
vendored under /root/reference nor installable here (no network).  This package restates, as dense torch code, the
published semantics of exactly the symbols the reference imports:

  kernel_spec.py:2-3   gpytorch.kernels.Kernel, RBFKernel
  kernel_gen.py:3      gpytorch.kernels.AdditiveKernel, ProductKernel, ScaleKernel
  LVAE.py:183-188      gpytorch.likelihoods.GaussianLikelihood, gpytorch.constraints.GreaterThan
  GP_def.py:8-21       gpytorch.models.ExactGP (state_dict container only)

so that the reference's own kernel_spec.py / kernel_gen.py / elbo_functions.py can be imported UNMODIFIED by
oracle/make_golden.py.  Nothing in the product package imports this.
"""
from . import constraints, kernels, likelihoods, means, models, distributions  # noqa: F401
