"""TEST INFRASTRUCTURE ONLY — CPU FP64 restatement (torch, CPU tensors) of the reference's GP-prior ELBO path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module, and only as the checker / the CPU baseline.  The product package never imports it.

Parity pin: the reference has no tests or golden vectors for this path.  This restatement is pinned
instead against the reference's OWN functions executed in the build container: `oracle/make_golden.py` imports
`/root/reference/{elbo_functions,kernel_gen,kernel_spec,GP_model}.py` unmodified (over the gpytorch stand-in in
`oracle/gpytorch_standin/`, because GPyTorch — an unpinned third-party dependency, README.MD:25 — is absent) and
stores their outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this file against those vectors.

Each function cites the reference lines it follows.  The op sequence and matmul shapes deliberately mirror the
reference (dense per-factor temporaries, explicit inverses through cholesky_solve, the mis-associated
`(Kxz Kzz^-1) m`) so that timing this port on host cores stands in for timing the reference.
"""
import math

import numpy as np
import torch

DT = torch.float64


# ----------------------------------------------------------------------------------------------------------------
# kernel structure  (kernel_gen.py:199-310, GP_model.py:146-236)
# ----------------------------------------------------------------------------------------------------------------
class Component:
    """outputscale[L] * prod(factors); factor = ('cat'|'bin', dim) or ('rbf', dim) with its own lengthscale[L]."""

    def __init__(self, factors, L):
        self.factors = list(factors)
        self.outputscale = torch.full((L,), math.log(2.0), dtype=DT)          # softplus(0), gpytorch ScaleKernel init
        self.lengthscales = {i: torch.full((L,), 2.5, dtype=DT)               # kernel_spec.py:68
                             for i, f in enumerate(self.factors) if f[0] == 'rbf'}

    def params(self):
        return [self.outputscale] + [self.lengthscales[i] for i in sorted(self.lengthscales)]


def parse_kernel_lists(L, cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel,
                       covariate_missing_val, id_covariate):
    """Component order cat, sqexp, bin, cat_int, bin_int; a component goes to K1 iff it is the id `cat_kernel` entry or a
    `cat_int_kernel` entry whose categorical covariate is the id (kernel_gen.py:225-308); a covariate listed in
    `covariate_missing_val` gains a `bin(mask)` factor right after it (kernel_gen.py:226-231, 246-251, 270-282)."""
    masks = {d['covariate']: d['mask'] for d in reversed(covariate_missing_val)}   # .index() -> first match wins

    def with_mask(kind, dim):
        return [(kind, dim)] + ([('bin', masks[dim])] if dim in masks else [])

    k0, k1 = [], []
    for d in cat_kernel:
        (k1 if d == id_covariate else k0).append(Component(with_mask('cat', d), L))
    for d in sqexp_kernel:
        k0.append(Component(with_mask('rbf', d), L))
    for d in bin_kernel:
        k0.append(Component(with_mask('bin', d), L))
    for e in cat_int_kernel:
        comp = Component(with_mask('cat', e['cat_covariate']) + with_mask('rbf', e['cont_covariate']), L)
        (k1 if e['cat_covariate'] == id_covariate else k0).append(comp)
    for e in bin_int_kernel:
        k0.append(Component(with_mask('bin', e['bin_covariate']) + with_mask('rbf', e['cont_covariate']), L))
    return k0, k1


def _factor(kind, dim, ls, x1, x2):
    a = x1[..., dim].unsqueeze(-1)
    b = x2[..., dim].unsqueeze(-2)
    if kind == 'cat':                                   # kernel_spec.py:31-32, GP_model.py:52-53
        return (a - b == 0).to(DT)
    if kind == 'bin':                                   # kernel_spec.py:22-23, GP_model.py:40-41
        return (a + b == 2).to(DT)
    s = ls.view(-1, 1, 1)                               # GP_model.py:79-85 (direct-difference SE), [L,1,1] broadcast
    return torch.exp(-((a - b) ** 2) / (2 * s ** 2))


def dense(components, x1, x2, L):
    """Additive kernel, latent dimension third from last: x [n,Q] | [L,n,Q] | [P,L,n,Q] -> [L,n1,n2] | [P,L,n1,n2]
    (gpytorch batch broadcasting of [L,1,1] parameters)."""
    lead = max(x1.dim(), x2.dim()) - 2
    out_shape = (L, x1.shape[-2], x2.shape[-2]) if lead <= 1 else (*x1.shape[:-3], L, x1.shape[-2], x2.shape[-2])
    total = torch.zeros(out_shape, dtype=DT)
    for comp in components:
        prod = None
        for i, (kind, dim) in enumerate(comp.factors):
            f = _factor(kind, dim, comp.lengthscales.get(i), x1, x2)
            prod = f if prod is None else prod * f
        total = total + comp.outputscale.view(L, 1, 1) * prod
    return total


# ----------------------------------------------------------------------------------------------------------------
# the bound  (elbo_functions.py:144-216 and 219-307)
# ----------------------------------------------------------------------------------------------------------------
def _chol_inv(A):
    Lc = torch.linalg.cholesky(A)
    eye = torch.eye(A.shape[-1], dtype=DT)
    return Lc, torch.cholesky_solve(eye, Lc)


def kld_fixed_T(k0, k1, noise, L, m, H, x, mu, log_v, z, P_tot, P_b, T, natural_gradient, eps):
    """elbo_functions.py:144-216.  noise [L]; x [P_b*T,Q] subject-major with exactly T rows/subject (no id check, 168)."""
    M = H.shape[-1]
    Q = x.shape[1]
    xs = x.reshape(P_b, T, Q).unsqueeze(1).expand(P_b, L, T, Q)                       # 168-169
    Kxz = dense(k0, x, z, L)                                                          # 171  [L,N,M]
    Kzz = dense(k0, z, z, L) + eps * torch.eye(M, dtype=DT)                            # 172,176
    K0s = dense(k0, xs, xs, L).transpose(0, 1)                                         # 173  [L,P,T,T]
    Bs = (dense(k1, xs, xs, L) + torch.eye(T, dtype=DT) * noise.view(L, 1, 1)).transpose(0, 1)   # 174
    Lz, Ki = _chol_inv(Kzz)                                                            # 177-178
    LB, Bi = _chol_inv(Bs)                                                             # 179-180
    Kxz_s = Kxz.reshape(L, P_b, T, M)
    BiK = Bi @ Kxz_s                                                                   # 183
    S = Kxz.transpose(1, 2) @ BiK.reshape(L, P_b * T, M)                               # 184
    LH, Hi = _chol_inv(H)                                                              # 185-186
    r = ((Kxz @ Ki) @ m).squeeze(-1) - mu.T                                            # 189 (mis-association kept)
    rs = r.reshape(L, P_b, T, 1)
    A = (rs.transpose(2, 3) @ Bi @ rs).sum()                                           # 190
    Bt = (torch.diagonal(Bi, dim1=-1, dim2=-2).reshape(L, -1) * torch.exp(log_v.T)).sum()   # 191
    C = 2 * torch.log(torch.diagonal(LB, dim1=-2, dim2=-1)).sum()                      # 192
    D = (Bi * K0s).sum() - (S * Ki).sum()                                              # 193
    G = Ki @ H @ Ki                                                                    # 194
    E = (G.transpose(-1, -2) * S).sum()                                                # 195
    F = log_v.sum()                                                                    # 196
    kl_qp = 0.5 * ((Ki * H.transpose(-1, -2)).sum() + (m * (Ki @ m)).sum() - L * M
                   + 2 * torch.log(torch.diagonal(Lz, dim1=-1, dim2=-2)).sum()
                   - 2 * torch.log(torch.diagonal(LH, dim1=-1, dim2=-2)).sum())        # 199-203
    kld = P_tot / P_b * 0.5 * (A + Bt + C + D + E - F) + kl_qp - L * P_tot * T / 2     # 204
    grad_m = grad_H = None
    if natural_gradient:                                                               # 208-214
        mus = mu.T.reshape(L, P_b, T, 1)
        ng1 = ((Ki.unsqueeze(1) @ Kxz_s.transpose(-1, -2)) @ (Bi @ mus)).sum(dim=1)
        Bm = Ki @ S @ Ki + Ki
        grad_m = -ng1 + Bm @ m
        grad_H = 0.5 * (-Hi + Bm)
    return kld, grad_m, grad_H


def kld_iter(k0, k1, noise, L, m, H, x, mu, log_v, z, P, P_b, N, natural_gradient, id_covariate, eps):
    """elbo_functions.py:219-307: ragged T, subjects = sorted unique ids, rows gathered by boolean mask (264-267)."""
    M = H.shape[-1]
    Kxz = dense(k0, x, z, L)
    Kzz = dense(k0, z, z, L) + eps * torch.eye(M, dtype=DT)
    Lz, Ki = _chol_inv(Kzz)
    LH, Hi = _chol_inv(H)
    r = (((Kxz @ Ki) @ m).squeeze(-1) - mu.T).unsqueeze(2)                              # 255
    G = Ki @ H @ Ki                                                                    # 256
    A = Bt = C = D = E = torch.zeros((), dtype=DT)
    ng1 = torch.zeros(L, M, 1, dtype=DT)
    S_all = torch.zeros(L, M, M, dtype=DT)
    ids = x[:, id_covariate]
    for s in torch.unique(ids).tolist():                                               # 264
        sel = ids == s
        xs = x[sel].unsqueeze(0).expand(L, -1, -1)
        Ts = xs.shape[1]
        K0s = dense(k0, xs, xs, L)
        Bs = dense(k1, xs, xs, L) + torch.eye(Ts, dtype=DT) * noise.view(L, 1, 1)       # 271
        LB, Bi = _chol_inv(Bs)
        Kp = Kxz[:, sel]
        S = torch.einsum('bik,bij,bjl->bkl', Kp, Bi, Kp)                               # 276
        rp = r[:, sel]
        A = A + torch.einsum('bji,bjk,bkl->b', rp, Bi, rp).sum()                       # 278
        Bt = Bt + (torch.diagonal(Bi, dim1=-1, dim2=-2).reshape(L, -1) * torch.exp(log_v[sel].T)).sum()
        C = C + 2 * torch.log(torch.diagonal(LB, dim1=-2, dim2=-1)).sum()
        D = D + (Bi * K0s).sum() - (S * Ki).sum()
        E = E + (G * S).sum()                                                          # 282 (no transpose here)
        if natural_gradient:
            ng1 = ng1 + Kp.transpose(-1, -2) @ (Bi @ mu[sel].T.unsqueeze(2))           # 287
            S_all = S_all + S
    F = log_v.sum()
    kl_qp = 0.5 * ((Ki * H.transpose(-1, -2)).sum() + (m * (Ki @ m)).sum() - L * M
                   + 2 * torch.log(torch.diagonal(Lz, dim1=-1, dim2=-2)).sum()
                   - 2 * torch.log(torch.diagonal(LH, dim1=-1, dim2=-2)).sum())
    kld = P / P_b * 0.5 * (A + Bt + C + D + E - F) + kl_qp - L * N / 2                  # 299
    grad_m = grad_H = None
    if natural_gradient:                                                               # 301-305
        Bm = Ki @ (S_all @ Ki) + Ki
        grad_m = -(Ki @ ng1) + Bm @ m
        grad_H = 0.5 * (-Hi + Bm)
    return kld, grad_m, grad_H


def ng_step(m, H, grad_m, grad_H, lr):
    """training.py:129-135 natural-gradient update of (m, H)."""
    _, iH = _chol_inv(H)
    iH_new = iH + lr * (grad_H + grad_H.transpose(-1, -2))
    _, H_new = _chol_inv(iH_new)
    m_new = H_new @ (iH @ m - lr * (grad_m - 2 * (grad_H @ m)))
    return m_new.detach(), H_new.detach()


# ----------------------------------------------------------------------------------------------------------------
# subject grouping / batch composition  (utils.py:40-113, training.py:69-75) — integer work, bit-exact
# ----------------------------------------------------------------------------------------------------------------
def subject_sampler_rows(perm, T):
    """utils.py:52-56: subject permutation -> row indices T*x .. T*x+T-1 in permutation order."""
    return [int(T * s + i) for s in perm for i in range(T)]


def fixed_T_batches(perm, T, subjects_per_batch):
    """BatchSampler(SubjectSampler, spb*T, drop_last=False) (training.py:73-75): last batch may be short."""
    rows = subject_sampler_rows(perm, T)
    bs = subjects_per_batch * T
    return [rows[i:i + bs] for i in range(0, len(rows), bs)]


def varying_T_index(ids):
    """utils.py:71-77: first-occurrence start index per distinct id, end = next start (assumes contiguous rows)."""
    ids = [int(v) for v in ids]
    seen, starts = set(), []
    for i, v in enumerate(ids):
        if v not in seen:
            seen.add(v)
            starts.append(i)
    return starts, starts[1:] + [len(ids)]


def varying_T_batches(ids, perm, subjects_per_batch):
    """utils.py:79-113: rows of shuffled subjects; a batch closes when spb distinct subjects are collected."""
    starts, ends = varying_T_index(ids)
    batches, cur, members = [], [], set()
    for s in perm:
        s = int(s)
        if s not in members:
            if len(members) == subjects_per_batch:
                batches.append(cur)
                cur, members = [], set()
            members.add(s)
        cur.extend(range(starts[s], ends[s]))
    batches.append(cur)
    return batches


def group_rows_by_subject(id_column):
    """elbo_functions.py:264-267: sorted unique ids and, per id, the row indices carrying it (any order of rows)."""
    idc = np.asarray(id_column, dtype=np.float64)
    uniq = np.unique(idc)
    return uniq, [np.nonzero(idc == u)[0] for u in uniq]
