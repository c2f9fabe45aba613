"""Drop-in for the reference's kernel_gen.py: the six structure lists -> additive kernel modules.

Block-indexing rule kept bit-exact (kernel_gen.py:225-308): component order is cat, sqexp, bin, cat_int, bin_int; a
component belongs to the id kernel K1 iff it is the `cat_kernel` entry on `id_covariate` or a `cat_int_kernel` entry
whose categorical covariate is `id_covariate`; everything else (all bin_int included) belongs to K0; a covariate listed
in `covariate_missing_val` is multiplied by `BinKernel(mask)`.  One deviation: the reference's batched generator raises
NameError (`Scalekernel`, kernel_gen.py:242) for a non-id `cat_kernel` entry; here that entry works.
"""
import torch

from .gp_kernels import AdditiveKernel, ProductKernel, ScaleKernel
from .kernel_spec import BinKernel, CatKernel, RbfKernel


def _assemble(latent_dim, cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel, covariate_missing_val,
              id_covariate):
    """Ordered list of (belongs_to_K1, ScaleKernel)."""
    batch = {} if latent_dim is None else {"batch_shape": torch.Size([latent_dim])}
    missing = [d['covariate'] for d in covariate_missing_val]

    def masked(kernel, covariate):
        if covariate in missing:
            mask_dim = covariate_missing_val[missing.index(covariate)]['mask']
            return kernel * BinKernel(active_dims=mask_dim, value=1)
        return kernel

    def rbf(dim):
        return RbfKernel(active_dims=dim, **batch) if batch else RbfKernel(active_dims=dim)

    out = []
    for d in cat_kernel:
        out.append((id_covariate is not None and d == id_covariate,
                    ScaleKernel(masked(CatKernel(active_dims=d), d), **batch)))
    for d in sqexp_kernel:
        out.append((False, ScaleKernel(masked(rbf(d), d), **batch)))
    for d in bin_kernel:
        out.append((False, ScaleKernel(masked(BinKernel(active_dims=d, value=1), d), **batch)))
    for e in cat_int_kernel:
        k1 = masked(CatKernel(active_dims=e['cat_covariate']), e['cat_covariate'])
        k2 = masked(rbf(e['cont_covariate']), e['cont_covariate'])
        out.append((id_covariate is not None and e['cat_covariate'] == id_covariate,
                    ScaleKernel(ProductKernel(k1, k2), **batch)))
    for e in bin_int_kernel:
        k1 = masked(BinKernel(active_dims=e['bin_covariate'], value=1), e['bin_covariate'])
        k2 = masked(rbf(e['cont_covariate']), e['cont_covariate'])
        out.append((False, ScaleKernel(ProductKernel(k1, k2), **batch)))
    return out


def _device():
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


def generate_kernel(cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel, covariate_missing_val):
    """One additive kernel over all components (kernel_gen.py:9-94)."""
    parts = _assemble(None, cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel,
                      covariate_missing_val, None)
    return AdditiveKernel(*[k for _, k in parts])


def generate_kernel_approx(cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel, covariate_missing_val,
                           id_covariate):
    """(K0 without the id covariate, K1 with it), un-batched (kernel_gen.py:97-197)."""
    parts = _assemble(None, cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel,
                      covariate_missing_val, id_covariate)
    return (AdditiveKernel(*[k for is1, k in parts if not is1]), AdditiveKernel(*[k for is1, k in parts if is1]))


def generate_kernel_batched(latent_dim, cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel,
                            covariate_missing_val, id_covariate):
    """(K0, K1) with batch_shape=[latent_dim] on every Scale/RBF kernel (kernel_gen.py:199-310)."""
    parts = _assemble(latent_dim, cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel,
                      covariate_missing_val, id_covariate)
    k0 = AdditiveKernel(*[k for is1, k in parts if not is1])
    k1 = AdditiveKernel(*[k for is1, k in parts if is1])
    return k0.to(_device()), k1.to(_device())
