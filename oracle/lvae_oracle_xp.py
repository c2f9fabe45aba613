"""TEST INFRASTRUCTURE ONLY — the forward quantities of the GP-prior bound in EXTENDED precision (numpy longdouble, x87
80-bit: unit round-off 5.4e-20, 2000 x finer than FP64), used as the yardstick when the reference's own FP64 result is not
accurate to the parity tolerance.

Why it exists: with inducing points drawn from the data rows, `Kzz + 1e-6 I` has cond 3e7 .. 1e9, and
`grad_m`, `grad_H` contain `Kzz^-1` twice.  Two correct FP64 implementations (LAPACK on the host, cuSOLVER on the device,
this repo's kernels) then differ by 1e-6 .. 1e-4 in those outputs — MORE than the 1e-6 parity tolerance — and the question
"which one is right" can only be answered against a more precise evaluation of the same formulas.  This module restates
elbo_functions.py:171-214 / 264-305 (same formulas as oracle/lvae_oracle.py, exact association of `Kxz Kzz^-1 m`) with
hand-written Cholesky / triangular inverse on longdouble arrays; no autograd, so only the outputs with closed forms:
kld per latent, grad_m, grad_H, d_mu, d_log_v.

Pinned by tests/test_oracle_golden.py: on the golden cases (well inside FP64's reach) it reproduces the reference-run
vectors to 1e-9.  Only tests/ and bench.py's parity check may import it.
"""
import numpy as np

XP = np.longdouble


def _chol(A):
    """Lower Cholesky factors of a batch [B, n, n] (right-looking, column by column)."""
    A = np.array(A, dtype=XP)
    B, n, _ = A.shape
    Lc = np.zeros_like(A)
    for j in range(n):
        s = A[:, j, j] - np.einsum('bk,bk->b', Lc[:, j, :j], Lc[:, j, :j])
        if np.any(s <= 0):
            raise np.linalg.LinAlgError("matrix not positive definite (extended-precision oracle)")
        d = np.sqrt(s)
        Lc[:, j, j] = d
        if j + 1 < n:
            Lc[:, j + 1:, j] = (A[:, j + 1:, j] - np.einsum('bik,bk->bi', Lc[:, j + 1:, :j], Lc[:, j, :j])) / d[:, None]
    return Lc


def _tri_inv(Lc):
    """Inverses of lower-triangular factors [B, n, n] by forward substitution."""
    B, n, _ = Lc.shape
    X = np.zeros_like(Lc)
    for i in range(n):
        rhs = -np.einsum('bk,bkj->bj', Lc[:, i, :i], X[:, :i, :])
        rhs[:, i] += 1
        X[:, i, :] = rhs / Lc[:, i, i][:, None]
    return X


def _spd_inv(A):
    Lc = _chol(A)
    Li = _tri_inv(Lc)
    return Lc, np.einsum('bki,bkj->bij', Li, Li)


def _factor(kind, dim, ls, x1, x2):
    a = x1[..., dim][..., :, None]
    b = x2[..., dim][..., None, :]
    if kind == 'cat':
        return (a - b == 0).astype(XP)
    if kind == 'bin':
        return (a + b == 2).astype(XP)
    return np.exp(-((a - b) ** 2) / (2 * ls * ls))


def _dense(components, x1, x2, l):
    """Additive kernel of latent l between covariate arrays [..., n1, Q] and [..., n2, Q] -> [..., n1, n2]."""
    total = None
    for comp in components:
        prod = None
        for i, (kind, dim) in enumerate(comp.factors):
            ls = XP(float(comp.lengthscales[i][l])) if i in comp.lengthscales else None
            f = _factor(kind, dim, ls, x1, x2)
            prod = f if prod is None else prod * f
        term = XP(float(comp.outputscale[l])) * prod
        total = term if total is None else total + term
    return total


def kld_forward(k0, k1, noise, latents, m, H, x, offsets, mu, log_v, z, scale, const_per_latent, eps):
    """Per-latent forward outputs in extended precision.  k0 / k1: oracle Components (their parameter tensors are read as
    floats), noise [L], m [L,M,1], H [L,M,M], x [N,Q], offsets [P+1] (subject rows contiguous), mu / log_v [N,L], z [L,M,Q]:
    torch CPU tensors or numpy arrays.  Returns dict of numpy float64 arrays restricted to `latents`:
    kld [n], grad_m [n,M], grad_H [n,M,M], d_mu [N,n], d_log_v [N,n]."""
    tonp = lambda t: np.asarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t)
    x = tonp(x).astype(XP)
    mu, log_v, z, m, H, noise = (tonp(t) for t in (mu, log_v, z, m, H, noise))
    offsets = np.asarray(offsets, dtype=np.int64)
    counts = np.diff(offsets)
    N, M = x.shape[0], z.shape[1]
    c = XP(scale) / 2
    out = dict(kld=[], grad_m=[], grad_H=[], d_mu=[], d_log_v=[])
    for l in latents:
        zl = z[l].astype(XP)
        ml = m[l].reshape(M).astype(XP)
        Hl = H[l].astype(XP)
        mul, lvl = mu[:, l].astype(XP), log_v[:, l].astype(XP)
        Kxz = _dense(k0, x, zl, l)
        Kzz = _dense(k0, zl, zl, l) + XP(eps) * np.eye(M, dtype=XP)
        Lz, Ki = _spd_inv(Kzz[None])
        Lz, Ki = Lz[0], Ki[0]
        LH, Hi = _spd_inv(Hl[None])
        LH, Hi = LH[0], Hi[0]
        a = Ki @ ml
        r = Kxz @ a - mul
        S = np.zeros((M, M), dtype=XP)
        ng1 = np.zeros(M, dtype=XP)
        A = Bt = C = D1 = XP(0)
        u = np.zeros(N, dtype=XP)
        bdiag = np.zeros(N, dtype=XP)
        for T in np.unique(counts):
            sel = np.nonzero(counts == T)[0]
            rows = (offsets[sel][:, None] + np.arange(T)[None, :])            # [n, T]
            xs = x[rows]                                                      # [n, T, Q]
            K0s = _dense(k0, xs, xs, l)
            Bs = _dense(k1, xs, xs, l) + XP(float(noise[l])) * np.eye(int(T), dtype=XP)
            LB, Bi = _spd_inv(Bs)
            Kp = Kxz[rows]                                                    # [n, T, M]
            V = np.einsum('ptu,puj->ptj', Bi, Kp)
            S += np.einsum('pti,ptj->ij', Kp, V)
            rp = r[rows]
            up = np.einsum('ptu,pu->pt', Bi, rp)
            u[rows] = up
            A += np.sum(rp * up)
            bd = np.einsum('ptt->pt', Bi)
            bdiag[rows] = bd
            Bt += np.sum(bd * np.exp(lvl[rows]))
            C += 2 * np.sum(np.log(np.einsum('ptt->pt', LB)))
            D1 += np.sum(Bi * K0s)
            ng1 += np.einsum('ptj,pt->j', V, mul[rows])
        D = D1 - np.sum(S * Ki)
        G = Ki @ Hl @ Ki
        E = np.sum(G.T * S)
        F = np.sum(lvl)
        kl = (np.sum(Ki * Hl.T) + ml @ (Ki @ ml) - M + 2 * np.sum(np.log(np.diag(Lz))) - 2 * np.sum(np.log(np.diag(LH)))) / 2
        kld = XP(scale) / 2 * (A + Bt + C + D + E - F) + kl - XP(const_per_latent)
        Bm = Ki @ S @ Ki + Ki
        out["kld"].append(kld)
        out["grad_m"].append(-(Ki @ ng1) + Bm @ ml)
        out["grad_H"].append((Bm - Hi) / 2)
        out["d_mu"].append(-2 * c * u)
        out["d_log_v"].append(c * (bdiag * np.exp(lvl) - 1))
    return dict(kld=np.array(out["kld"], dtype=np.float64), grad_m=np.array(out["grad_m"], dtype=np.float64),
                grad_H=np.array(out["grad_H"], dtype=np.float64), d_mu=np.array(out["d_mu"], dtype=np.float64).T,
                d_log_v=np.array(out["d_log_v"], dtype=np.float64).T)
