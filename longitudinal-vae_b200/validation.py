"""Drop-in for the reference's validation.validation_dubo (validation.py:8-68) on the GPU ops (differentiable, see
elbo_functions._low_rank_terms)."""
from .elbo_functions import _dubo_per_latent


def validation_dubo(latent_dim, covar_module0, covar_module1, likelihood, train_xt, m, log_v, z, P, T, eps):
    """Sum over latent dimensions of the deviance upper bound (DUBO) with batched kernel modules; returns a [1] tensor."""
    return _dubo_per_latent(latent_dim, covar_module0, covar_module1, likelihood, train_xt, m, log_v, z, P, T, eps).sum().reshape(1)
