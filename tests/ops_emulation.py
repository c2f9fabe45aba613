"""TEST INFRASTRUCTURE ONLY — torch-CPU stand-ins for the C-ABI ops that diff_ops.py composes, so the autograd plumbing of
the non-minibatch bounds (which adjoint feeds which op, shapes, transposes, signs) can be checked against the reference's
golden gradients on a machine without a GPU.  The CUDA kernels themselves are checked by the `-m gpu` tests; nothing in
the package imports this file."""
import contextlib

import torch

from lvae_b200._lib import CAT

F64 = torch.float64


def _rng(structure, which):
    return {"k0": (0, structure.n_comp0), "k1": (structure.n_comp0, structure.n_comp), "all": (0, structure.n_comp)}[which]


def _dense(structure, ls, os_, x1, x2, which, diag):
    L = os_.shape[1]
    B = max(L, x1.shape[0] if x1.dim() == 3 else 1, x2.shape[0] if x2.dim() == 3 else 1)
    a = x1.to(F64) if x1.dim() == 3 else x1.to(F64).unsqueeze(0).expand(B, -1, -1)
    b = x2.to(F64) if x2.dim() == 3 else x2.to(F64).unsqueeze(0).expand(B, -1, -1)
    n1, n2 = a.shape[1], b.shape[1]
    dev = a.device
    lat = torch.arange(B, device=dev) % L
    out = torch.zeros(B, n1, n2, dtype=F64, device=dev)
    lo, hi = _rng(structure, which)
    for c in range(lo, hi):
        row = [int(v) for v in structure.table[c]]
        f = torch.ones(B, n1, n2, dtype=F64, device=dev)
        for i in range(row[2]):
            ty, dm = row[3 + 2 * i], row[4 + 2 * i]
            u, v = a[:, :, dm].unsqueeze(2), b[:, :, dm].unsqueeze(1)
            f = f * ((u - v == 0) if ty == CAT else (u + v == 2)).to(F64)
        if row[0] >= 0:
            d = a[:, :, row[0]].unsqueeze(2) - b[:, :, row[0]].unsqueeze(1)
            ell = ls[row[1]][lat].view(B, 1, 1)
            f = f * torch.exp(-d * d / (2 * ell * ell))
        out = out + os_[c][lat].view(B, 1, 1) * f
    if diag is not None:
        out = out + diag[lat].view(B, 1, 1) * torch.eye(n1, n2, dtype=F64, device=dev)
    return out


def kernel_dense(structure, lengthscale, outputscale, x1, x2, which="all", diag_add=None):
    with torch.no_grad():
        return _dense(structure, lengthscale, outputscale, x1, x2, which, diag_add)


def _blocks(structure, ls, os_, x, offsets, which, diag):
    off = offsets.tolist()
    L, P = os_.shape[1], len(off) - 1
    T = off[1] - off[0] if P else 0
    if P and all(off[p + 1] - off[p] == T for p in range(P)):         # equal T: one stacked evaluation, matrix p*L + l
        xs = x.reshape(P, 1, T, -1).expand(P, L, T, x.shape[-1]).reshape(P * L, T, -1)
        return _dense(structure, ls, os_, xs, xs, which, diag).view(P, L, T * T).permute(1, 0, 2).reshape(L, P * T * T)
    parts = []
    for p in range(len(off) - 1):
        xp = x[off[p]:off[p + 1]]
        parts.append(_dense(structure, ls, os_, xp, xp, which, diag).reshape(os_.shape[1], -1))
    return torch.cat(parts, dim=1)


def kernel_blocks(structure, lengthscale, outputscale, x, offsets_dev, sum_T2, which="k0", diag_add=None):
    with torch.no_grad():
        out = _blocks(structure, lengthscale, outputscale, x, offsets_dev, which, diag_add)
    assert out.shape[1] == sum_T2
    return out


def _bwd(fn, ls, os_, diag, g):
    ls = ls.detach().clone().requires_grad_(True)
    os_ = os_.detach().clone().requires_grad_(True)
    dg = torch.zeros(os_.shape[1], dtype=F64, device=os_.device, requires_grad=True) if diag else None
    with torch.enable_grad():
        out = fn(ls, os_, dg)
        grads = torch.autograd.grad(out, [ls, os_] + ([dg] if diag else []), g, allow_unused=True)
    z = lambda t, like: torch.zeros_like(like) if t is None else t
    return z(grads[0], ls), z(grads[1], os_), (z(grads[2], dg) if diag else None)


def kernel_dense_bwd(structure, lengthscale, outputscale, x1, x2, grad_out, which="all", want_diag=False):
    return _bwd(lambda l, o, d: _dense(structure, l, o, x1, x2, which, d), lengthscale, outputscale, want_diag, grad_out)


def kernel_blocks_bwd(structure, lengthscale, outputscale, x, offsets_dev, grad_out, which="k0", want_diag=False):
    return _bwd(lambda l, o, d: _blocks(structure, l, o, x, offsets_dev, which, d), lengthscale, outputscale, want_diag,
                grad_out)


def potrf_batched(A, check_info=True):
    return torch.linalg.cholesky(A.detach())


def potri_batched(Lc):
    return torch.cholesky_inverse(Lc.detach())


def gemm_batched(A, B, trans_a=False, trans_b=False, alpha=1.0, beta=0.0, C=None, flags=0):
    out = alpha * torch.bmm(A.detach().transpose(1, 2) if trans_a else A.detach(),
                            B.detach().transpose(1, 2) if trans_b else B.detach())
    return out if C is None or beta == 0.0 else out + beta * C


@contextlib.contextmanager
def emulated_ops():
    """Swap the C-ABI wrappers of lvae_b200.ops (and the CUDA guard of the bounds) for the stand-ins above."""
    import lvae_b200.elbo_functions as EF
    from lvae_b200 import ops
    names = ["kernel_dense", "kernel_blocks", "kernel_dense_bwd", "kernel_blocks_bwd", "potrf_batched", "potri_batched",
             "gemm_batched"]
    saved = {n: getattr(ops, n) for n in names}
    guard = EF._need_cuda
    try:
        for n in names:
            setattr(ops, n, globals()[n])
        EF._need_cuda = lambda x: None
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
        EF._need_cuda = guard
