"""Synthetic Health-MNIST-shaped inputs for the GP-prior ELBO path (no data set ships; there is no network).

Mirrors the covariate layout the reference loads in `dataset_def.py:172-219` (label columns
`[time_age, disease_time, subject, gender, disease, location]`, float32-rounded, id column 2) and the time grids of
`Health_MNIST_generate.py:89-90,118-154` (time_age 0..T-1, disease_time = t-9 for diseased subjects, NaN->0 otherwise).
Inducing points are M data rows shared by every latent (`LVAE.py:199-203`), `m ~ N(0,1)`, `H = (A/10)(A/10)^T`
(`LVAE.py:222-226`).  Used by tests, bench.py and the golden-vector generator; pure numpy + torch CPU.
"""
from dataclasses import dataclass, field

import numpy as np
import torch

Q_COVARIATES = 6
ID_COVARIATE = 2

# kernel-structure lists exactly as the reference's config passes them (parse_model_args.py:74-79)
_SAMPLE_CAT_INT = [{'cont_covariate': 0, 'cat_covariate': 2}, {'cont_covariate': 0, 'cat_covariate': 3},
                   {'cont_covariate': 1, 'cat_covariate': 4}]

CONFIGS = {
    # name: (P, T or (Tmin,Tmax), L, M, kernel lists)
    "cfg1": dict(P=100, T=20, L=32, M=60, cat_kernel=[], bin_kernel=[], sqexp_kernel=[0],
                 cat_int_kernel=[{'cont_covariate': 0, 'cat_covariate': 2}], bin_int_kernel=[],
                 covariate_missing_val=[]),
    "cfg2": dict(P=1000, T=20, L=32, M=60, cat_kernel=[2], bin_kernel=[], sqexp_kernel=[0],
                 cat_int_kernel=_SAMPLE_CAT_INT, bin_int_kernel=[], covariate_missing_val=[]),
    "cfg3": dict(P=1000, T=20, L=64, M=256, cat_kernel=[2], bin_kernel=[], sqexp_kernel=[0],
                 cat_int_kernel=_SAMPLE_CAT_INT, bin_int_kernel=[], covariate_missing_val=[]),
    "cfg4": dict(P=20000, T=(5, 40), L=32, M=60, cat_kernel=[2], bin_kernel=[5], sqexp_kernel=[0],
                 cat_int_kernel=_SAMPLE_CAT_INT, bin_int_kernel=[], covariate_missing_val=[]),
    "cfg5": dict(P=200000, T=20, L=64, M=128, cat_kernel=[2], bin_kernel=[], sqexp_kernel=[0],
                 cat_int_kernel=_SAMPLE_CAT_INT, bin_int_kernel=[], covariate_missing_val=[]),
}
KERNEL_LIST_KEYS = ("cat_kernel", "bin_kernel", "sqexp_kernel", "cat_int_kernel", "bin_int_kernel",
                    "covariate_missing_val")


def kernel_lists(cfg):
    """The six structure lists of a config, in `generate_kernel_batched` argument order (after latent_dim)."""
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    return {k: c[k] for k in KERNEL_LIST_KEYS}


@dataclass
class SynthBatch:
    x: torch.Tensor          # [N, Q] float64 covariates, subject-major contiguous rows
    offsets: np.ndarray      # [P+1] int64 CSR row offsets per subject
    mu: torch.Tensor         # [N, L]
    log_v: torch.Tensor      # [N, L]
    z: torch.Tensor          # [L, M, Q]
    m: torch.Tensor          # [L, M, 1]
    H: torch.Tensor          # [L, M, M] SPD
    P: int
    T: object
    L: int
    M: int
    lists: dict = field(default_factory=dict)

    @property
    def N(self):
        return int(self.x.shape[0])


def covariates(P, T, rng, first_subject=0):
    """Covariate matrix for subjects first_subject..first_subject+P-1; T int or (lo, hi) inclusive range."""
    if isinstance(T, (tuple, list)):
        lens = rng.integers(T[0], T[1] + 1, size=P)
    else:
        lens = np.full(P, int(T), dtype=np.int64)
    offsets = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    disease = rng.integers(0, 2, size=P)
    location = rng.integers(0, 2, size=P)
    x = np.zeros((int(offsets[-1]), Q_COVARIATES), dtype=np.float32)
    subj = np.repeat(np.arange(P), lens)
    t = np.arange(int(offsets[-1])) - np.repeat(offsets[:-1], lens)
    x[:, 0] = t
    x[:, 1] = np.where(disease[subj] == 1, t - 9, 0)
    x[:, 2] = subj + first_subject
    x[:, 3] = (subj + first_subject) % 2
    x[:, 4] = disease[subj]
    x[:, 5] = location[subj]
    return torch.from_numpy(x).double(), offsets


def make_batch(cfg, seed=None, P=None, L=None, M=None, T=None, first_subject=0):
    """Seeded synthetic problem for config `cfg` ("cfg1".."cfg5"), optionally shrunk (P/L/M/T overrides)."""
    c = dict(CONFIGS[cfg])
    idx = int(cfg[-1])
    seed = 1234 + idx if seed is None else seed
    P = c["P"] if P is None else P
    L = c["L"] if L is None else L
    M = c["M"] if M is None else M
    T = c["T"] if T is None else T
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    x, offsets = covariates(P, T, rng, first_subject)
    N = x.shape[0]
    mu = torch.randn(N, L, generator=g, dtype=torch.float64)
    log_v = -3.0 * torch.rand(N, L, generator=g, dtype=torch.float64)
    rows = np.sort(rng.choice(N, size=M, replace=False))
    z = x[rows].unsqueeze(0).repeat(L, 1, 1).contiguous()
    m = torch.randn(L, M, 1, generator=g, dtype=torch.float64)
    A = torch.randn(L, M, M, generator=g, dtype=torch.float64) / 10
    H = A @ A.transpose(-1, -2)
    return SynthBatch(x=x, offsets=offsets, mu=mu, log_v=log_v, z=z, m=m, H=H, P=P, T=T, L=L, M=M,
                      lists=kernel_lists(c))


def perturbed_hypers(n_lengthscale, n_outputscale, L, seed, noise_trainable=False):
    """Perturbed hyper-parameter set of SURVEY 8(d): l = 2.5 e^{0.3 xi}, s2 = ln2 e^{0.3 xi}, so latents differ."""
    g = torch.Generator().manual_seed(seed + 77)
    ls = 2.5 * torch.exp(0.3 * torch.randn(n_lengthscale, L, generator=g, dtype=torch.float64))
    os_ = np.log(2.0) * torch.exp(0.3 * torch.randn(n_outputscale, L, generator=g, dtype=torch.float64))
    noise = torch.ones(L, dtype=torch.float64)
    if noise_trainable:
        noise = torch.exp(0.2 * torch.randn(L, generator=g, dtype=torch.float64))
    return ls, os_, noise
