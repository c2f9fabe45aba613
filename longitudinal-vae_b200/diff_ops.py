"""Differentiable building blocks over the C ABI, for the paths the reference trains through plain autograd
(deviance_upper_bound / elbo, elbo_functions.py:36-142; `covar_module(x1, x2).evaluate()` under a loss, gpytorch's lazy
kernels).  Each Function runs a CUDA op of liblvae_b200.so forward and the matching adjoint backward; nothing here is a
torch re-implementation of the op itself.  The Hensman minibatch bound does NOT use these: it has one fused forward +
backward (elbo_functions._KldBound).

Gradients flow to the kernel hyper-parameters and the likelihood noise (tables [rows, L] produced by spec.build_structure),
never to the covariates or the inducing inputs — both are constants in the reference (LVAE.py:204-208).
"""
import torch

from . import ops


class KernelDense(torch.autograd.Function):
    """lvae_kernel_dense_f64 / lvae_kernel_dense_bwd_f64."""

    @staticmethod
    def forward(ctx, structure, which, x1, x2, lengthscale, outputscale, diag_add):
        ctx.structure, ctx.which, ctx.has_diag = structure, which, diag_add is not None
        ctx.save_for_backward(x1, x2, lengthscale, outputscale)
        return ops.kernel_dense(structure, lengthscale, outputscale, x1, x2, which, diag_add=diag_add)

    @staticmethod
    def backward(ctx, g):
        x1, x2, ls, os_ = ctx.saved_tensors
        d_ls, d_os, d_diag = ops.kernel_dense_bwd(ctx.structure, ls, os_, x1, x2, g, ctx.which, want_diag=ctx.has_diag)
        return None, None, None, None, d_ls, d_os, d_diag


class KernelBlocks(torch.autograd.Function):
    """lvae_kernel_blocks_f64 / lvae_kernel_blocks_bwd_f64; output flat [L, sum_T2]."""

    @staticmethod
    def forward(ctx, structure, which, x, offsets_dev, sum_T2, lengthscale, outputscale, diag_add):
        ctx.structure, ctx.which, ctx.has_diag = structure, which, diag_add is not None
        ctx.save_for_backward(x, offsets_dev, lengthscale, outputscale)
        return ops.kernel_blocks(structure, lengthscale, outputscale, x, offsets_dev, sum_T2, which, diag_add=diag_add)

    @staticmethod
    def backward(ctx, g):
        x, off, ls, os_ = ctx.saved_tensors
        d_ls, d_os, d_diag = ops.kernel_blocks_bwd(ctx.structure, ls, os_, x, off, g, ctx.which, want_diag=ctx.has_diag)
        return None, None, None, None, None, d_ls, d_os, d_diag


def _mm(A, B, ta=False, tb=False, flags=0):
    """Every product goes to lvae_gemm_batched_f64: the tiled DMMA GEMM for the per-latent matrices, its one-CTA-per-matrix
    kernel for the stacks of thousands of tiny per-subject blocks (T x T, T <= 40, one per subject and latent)."""
    m, k = (A.shape[2], A.shape[1]) if ta else (A.shape[1], A.shape[2])
    n = B.shape[1] if tb else B.shape[2]
    if min(m, n, k) > 0 and A.shape[0] > 0:
        return ops.gemm_batched(A, B, trans_a=ta, trans_b=tb, flags=flags)
    return torch.zeros(A.shape[0], m, n, dtype=A.dtype, device=A.device)


class Gemm(torch.autograd.Function):
    """C[b] = op(A[b]) op(B[b]) (lvae_gemm_batched_f64), with the two adjoint products on the same kernel.
    flags=3 computes the lower triangle and mirrors it — only for products known to be symmetric."""

    @staticmethod
    def forward(ctx, A, B, ta, tb, flags):
        if ta and tb:
            raise NotImplementedError("Gemm: op(A)=A^T with op(B)=B^T is not used on this path")
        ctx.ta, ctx.tb = ta, tb
        ctx.save_for_backward(A, B)
        return _mm(A.detach(), B.detach(), ta, tb, flags)

    @staticmethod
    def backward(ctx, G):
        A, B = ctx.saved_tensors
        G = G.contiguous()
        dA = dB = None
        if not ctx.ta and not ctx.tb:                        # C = A B
            if ctx.needs_input_grad[0]:
                dA = _mm(G, B, False, True)
            if ctx.needs_input_grad[1]:
                dB = _mm(A, G, True, False)
        elif ctx.ta:                                         # C = A^T B
            if ctx.needs_input_grad[0]:
                dA = _mm(B, G, False, True)
            if ctx.needs_input_grad[1]:
                dB = _mm(A, G, False, False)
        else:                                                # C = A B^T
            if ctx.needs_input_grad[0]:
                dA = _mm(G, B, False, False)
            if ctx.needs_input_grad[1]:
                dB = _mm(G, A, True, False)
        return dA, dB, None, None, None


def gemm(A, B, ta=False, tb=False, flags=0):
    return Gemm.apply(A, B, ta, tb, flags)


class SpdInverse(torch.autograd.Function):
    """(A^-1, log det A) of a batch of SPD matrices [batch, n, n] from lvae_potrf_batched_f64 + lvae_potri_batched_f64 — the
    `torch.cholesky` + `cholesky_solve(I, L)` / `torch.solve` pairs of elbo_functions.py:48-66, 104-124.
    Adjoint: dA = -A^-1 (dA^-1) A^-1 + (dlogdet) A^-1."""

    @staticmethod
    def forward(ctx, A):
        ctx.set_materialize_grads(False)
        Lc = ops.potrf_batched(A)
        Ai = ops.potri_batched(Lc)
        logdet = 2.0 * torch.log(torch.diagonal(Lc, dim1=-2, dim2=-1)).sum(-1)
        ctx.save_for_backward(Ai)
        return Ai, logdet

    @staticmethod
    def backward(ctx, g_inv, g_logdet):
        (Ai,) = ctx.saved_tensors
        dA = None
        if g_inv is not None:
            dA = -_mm(_mm(Ai, g_inv.contiguous()), Ai)
        if g_logdet is not None:
            t = g_logdet.reshape(-1, 1, 1) * Ai
            dA = t if dA is None else dA + t
        return dA


def spd_inverse(A):
    return SpdInverse.apply(A)
