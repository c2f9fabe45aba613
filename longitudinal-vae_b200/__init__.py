"""lvae_b200 — B200-native (hand-written sm_100a CUDA, FP64) GP-prior ELBO path of the Longitudinal VAE.

Drop-in mirror of the reference's Python API for that path:

    lvae_b200.elbo_functions.minibatch_KLD_upper_bound / minibatch_KLD_upper_bound_iter   (elbo_functions.py:144-307)
    lvae_b200.kernel_gen.generate_kernel_batched / generate_kernel / generate_kernel_approx (kernel_gen.py)
    lvae_b200.kernel_spec.{BinKernel, CatKernel, CatKernelMod, RbfKernel}                  (kernel_spec.py)
    lvae_b200.GP_model.{Likelihoods, BinKernel, ..., generate_kernel_batched}              (GP_model.py)
    lvae_b200.training.{natural_gradient_step, hensman_training}                           (training.py:15-237)
    lvae_b200.utils.{SubjectSampler, VaryingLengthSubjectSampler, VaryingLengthBatchSampler, HensmanDataLoader}

The directory is named `longitudinal-vae_b200/` (not importable as-is); `import lvae_b200` resolves here through the
alias package at the repository root.  All arithmetic runs in liblvae_b200.so (C ABI in include/lvae_b200.h); there
is no CPU fallback — calling an op without the library or without CUDA raises.
"""
__version__ = "0.1.0"
