"""Drop-in for the sampler half of the reference's utils.py (utils.py:9-113): subject-wise batching for the Hensman loop.

Semantics kept bit-exact under the same numpy RNG state (tests/golden/samplers.npz):
  SubjectSampler                 np.random.shuffle(arange(P)), then rows T*s .. T*s+T-1 of each subject in that order
  VaryingLengthSubjectSampler    ids = int(label[id_covariate]); a subject's rows run from the first occurrence of its id to
                                 the first occurrence of the next new id (rows assumed contiguous); shuffled subjects
  VaryingLengthBatchSampler      a batch closes when `batch_size` distinct subjects have been collected
  HensmanDataLoader              one persistent iterator over an endlessly repeating batch sampler; len = batches/epoch
"""
import numpy as np
import torch
from torch.utils.data.sampler import BatchSampler, Sampler


class _RepeatSampler:
    """Wraps a (batch) sampler so that iterating never ends."""

    def __init__(self, sampler):
        self.sampler = sampler

    def __iter__(self):
        while True:
            for item in self.sampler:
                yield item


class HensmanDataLoader(torch.utils.data.dataloader.DataLoader):
    def __init__(self, dataset, batch_sampler, num_workers):
        super().__init__(dataset, batch_sampler=_RepeatSampler(batch_sampler), num_workers=num_workers)
        self.iterator = super().__iter__()

    def __len__(self):
        return len(self.batch_sampler.sampler)

    def __iter__(self):
        for _ in range(len(self)):
            yield next(self.iterator)


class SubjectSampler(Sampler):
    def __init__(self, data_source, P, T):
        self.data_source, self.P, self.T = data_source, P, T

    def __iter__(self):
        order = np.arange(self.P)
        np.random.shuffle(order)
        rows = (order[:, None] * self.T + np.arange(self.T)[None, :]).reshape(-1)
        return iter(rows.tolist())

    def __len__(self):
        return len(self.data_source)


class VaryingLengthSubjectSampler(Sampler):
    def __init__(self, data_source, id_covariate):
        self.data_source, self.id_covariate = data_source, id_covariate
        ids = [int(sample['label'][id_covariate].item()) for sample in data_source]
        first = {}
        for row, v in enumerate(ids):
            first.setdefault(v, row)
        self.P = len(first)
        self.start_indices = list(first.values())                       # order of first appearance
        self.end_indices = self.start_indices[1:] + [len(data_source)]

    def __iter__(self):
        order = np.arange(self.P)
        np.random.shuffle(order)
        return iter([(row, int(s)) for s in order for row in range(self.start_indices[s], self.end_indices[s])])

    def __len__(self):
        return self.P


class VaryingLengthBatchSampler(BatchSampler):
    def __init__(self, sampler, batch_size):
        super().__init__(sampler, batch_size, False)
        assert isinstance(sampler, VaryingLengthSubjectSampler)
        self.sampler, self.batch_size = sampler, batch_size

    def __iter__(self):
        rows, members = [], set()
        for row, subject in self.sampler:
            if subject not in members:
                if len(members) == self.batch_size:
                    yield rows
                    rows, members = [], set()
                members.add(subject)
            rows.append(row)
        yield rows


# ---------------------------------------------------------------------------------------------------------------
# GP posterior-mean prediction (utils.py:115-345), SURVEY 8f-2.  Same call signatures as the reference; forward only.
# Kernel matrices come from lvae_kernel_dense_f64 / lvae_kernel_blocks_f64, the per-subject and M x M factorisations and
# explicit inverses from lvae_potrf_batched_f64 / lvae_potri_batched_f64, S = K0zx B^-1 K0xz from the batched DMMA GEMM;
# every product, matrix-vector ones included, is lvae_gemm_batched_f64 (no cuBLAS on this path).  No CPU fallback.
# ---------------------------------------------------------------------------------------------------------------
def _spd_inverse(A):
    from . import ops
    return ops.potri_batched(ops.potrf_batched(A))


def _predict_one(L, covar_module0, covar_module1, likelihoods, x, test_x, mu, z, id_covariate, eps):
    """x [N,Q] prediction (training) rows, mu [N,L], z [L,M,Q] | [M,Q]; returns Z_pred [N*, L]."""
    from . import ops
    from . import elbo_functions as EF
    from .elbo_functions import _noise_of, _structure_of, group_by_subject
    EF._need_cuda(x)                                                              # no CPU fallback
    f64 = torch.float64
    x, test_x, mu = x.to(f64), test_x.to(f64), mu.to(f64).reshape(x.shape[0], L)
    z = z.to(f64)
    if z.dim() == 2:
        z = z.unsqueeze(0).expand(L, -1, -1)
    z = z.contiguous()
    order, offsets, T_max, sum_T2 = group_by_subject(x[:, id_covariate])      # utils.py:160-163 (sorted unique ids)
    if order is not None:
        x, mu = x[order], mu[order]
    N, M = x.shape[0], z.shape[1]
    st, ls, os_ = _structure_of(covar_module0, covar_module1, L, x.device)
    ls, os_ = ls.detach(), os_.detach()
    noise = _noise_of(likelihoods, L, f64, x.device).detach().contiguous()
    K0xz = ops.kernel_dense(st, ls, os_, x, z, "k0")                              # [L,N,M]
    K0zz = ops.kernel_dense(st, ls, os_, z, z, "k0") + eps * torch.eye(M, dtype=f64, device=x.device)
    K0Xz = ops.kernel_dense(st, ls, os_, test_x, z, "k0")                         # [L,N*,M]
    # per-subject B_p = K1 + noise I, explicit inverses (utils.py:165-176), grouped by the number of rows per subject
    blocks = ops.kernel_blocks(st, ls, os_, x, offsets, sum_T2, "k1", diag_add=noise)      # flat [L, sum T^2]
    off = offsets.to(torch.int64).cpu()
    Ts = (off[1:] - off[:-1])
    off2 = torch.zeros_like(off)
    off2[1:] = torch.cumsum(Ts * Ts, 0)
    muL = mu.t().contiguous()                                                     # [L,N]
    iB_K0xz = torch.empty_like(K0xz)
    iB_mu = torch.empty(L, N, dtype=f64, device=x.device)
    groups = []
    for T in torch.unique(Ts).tolist():
        subj = torch.nonzero(Ts == T).reshape(-1)
        n = subj.numel()
        bidx = (off2[subj].unsqueeze(1) + torch.arange(T * T)).reshape(-1).to(x.device)       # block entries
        ridx = (off[subj].unsqueeze(1) + torch.arange(T)).reshape(-1).to(x.device)            # rows
        B = blocks[:, bidx].reshape(L * n, T, T)
        iB = _spd_inverse(B)                                                       # [L*n,T,T]
        Kx = K0xz[:, ridx].reshape(L * n, T, M)
        iB_K0xz[:, ridx] = ops.gemm_batched(iB, Kx).reshape(L, n * T, M)
        iB_mu[:, ridx] = ops.gemm_batched(iB, muL[:, ridx].reshape(L * n, T, 1)).reshape(L, n * T)
        groups.append((ridx, iB, n, T))
    H = K0zz + ops.gemm_batched(K0xz, iB_K0xz, trans_a=True)                       # K0zz + K0zx B^-1 K0xz
    H = 0.5 * (H + H.transpose(1, 2))
    rhs = ops.gemm_batched(K0xz, iB_mu.unsqueeze(2), trans_a=True)                 # [L,M,1]
    t1 = ops.gemm_batched(K0xz, ops.gemm_batched(_spd_inverse(H), rhs)).squeeze(2)   # K0xz H^-1 K0zx B^-1 mu   [L,N]
    mu_tilde = iB_mu.clone()
    for ridx, iB, n, T in groups:
        mu_tilde[:, ridx] -= ops.gemm_batched(iB, t1[:, ridx].reshape(L * n, T, 1)).reshape(L, n * T)
    pred0 = ops.gemm_batched(K0Xz, ops.gemm_batched(_spd_inverse(K0zz),
                                                    ops.gemm_batched(K0xz, mu_tilde.unsqueeze(2), trans_a=True)))
    # id-dependent part: K1(X*, X[mask]) mu_tilde[mask] over the subjects that occur in the test set (utils.py:187-206)
    test_subjects = torch.unique(test_x[:, id_covariate])
    mask = torch.isin(x[:, id_covariate], test_subjects)
    pred1 = torch.zeros_like(pred0)
    if bool(mask.any()):
        K1Xx = ops.kernel_dense(st, ls, os_, test_x, x[mask].contiguous(), "k1")   # [L,N*,Nmask]
        pred1 = ops.gemm_batched(K1Xx, mu_tilde[:, mask].unsqueeze(2).contiguous())
    return (pred0 + pred1).squeeze(2).t().contiguous()


def _predict_any(latent_dim, covar_module0, covar_module1, likelihoods, prediction_x, test_x, mu, zt_list, id_covariate, eps):
    if isinstance(covar_module0, (list, tuple)):                                   # one un-batched module per latent dimension
        cols = [_predict_one(1, covar_module0[i], covar_module1[i], likelihoods[i], prediction_x, test_x, mu[:, i:i + 1],
                             zt_list[i], id_covariate, eps) for i in range(latent_dim)]
        return torch.cat(cols, dim=1)
    return _predict_one(latent_dim, covar_module0, covar_module1, likelihoods, prediction_x, test_x, mu, zt_list,
                        id_covariate, eps)


def batch_predict_varying_T(latent_dim, covar_module0, covar_module1, likelihoods, prediction_x, test_x, mu, zt_list,
                            id_covariate, eps):
    """GP posterior-mean prediction for subjects with varying numbers of rows (utils.py:115-208).  Z_pred [N*, L]."""
    return _predict_any(latent_dim, covar_module0, covar_module1, likelihoods, prediction_x, test_x, mu, zt_list,
                        id_covariate, eps)


def batch_predict(latent_dim, covar_module0, covar_module1, likelihoods, prediction_x, test_x, mu, zt_list, P, T,
                  id_covariate, eps):
    """GP posterior-mean prediction, exactly T rows per subject in subject-major order (utils.py:210-299).  The reference
    reshapes by position ([P, T, Q], no id check).  Here the rows are grouped by the id column (sorted unique ids, like
    batch_predict_varying_T): identical for the input the reference is written for (every block of T consecutive rows is
    one subject, ids distinct between blocks); it differs only where the reference's positional reshape would mix subjects
    — an id recurring in non-adjacent blocks or a block holding several ids — which this function treats by id."""
    if prediction_x.shape[0] != P * T:
        raise RuntimeError(f"shape '[{P}, {T}, {prediction_x.shape[1]}]' is invalid for input of size {prediction_x.numel()}")
    return _predict_any(latent_dim, covar_module0, covar_module1, likelihoods, prediction_x, test_x, mu, zt_list,
                        id_covariate, eps)


def predict(covar_module0, covar_module1, likelihood, train_xt, test_x, mu, z, P, T, id_covariate, eps):
    """Single-latent prediction helper (utils.py:301-345): un-batched kernels, mu [N], z [M,Q]; returns Z_pred [N*]."""
    return _predict_one(1, covar_module0, covar_module1, likelihood, train_xt, test_x, mu.reshape(-1, 1), z, id_covariate,
                        eps).reshape(-1)
