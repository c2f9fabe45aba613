"""Full-size GPU parity (-m gpu): the CUDA path on the problems bench.py TIMES — BASELINE configs[1..4] at the bench's sizes —
against the oracle port of the reference (bench.check_parity, the very check the bench line's `parity` object records: all
latents against the oracle with stock torch CUDA ops as a gross-error net, four latents x ALL subjects against the oracle on
the host — LAPACK, the reference's own arithmetic — at 1e-6 + twice the oracle's measured input-rounding sensitivity).  Sizes: cfg2 1000 subjects x 20 rows, L=32, M=60; cfg3 1000 x 20, L=64, M=256; cfg4
2000 ragged subjects (5..40 rows), L=32, M=60; cfg5 2000 x 20, L=64, M=128.  Tolerance 1e-6, max-norm relative per output
tensor (kld, grad_m, grad_H, d_mu, d_log_v, the hyper-gradient vector, and (m, H) after the natural-gradient update).
Size-independent properties at the same sizes: fixed-T == iter on regular input (SURVEY 4 identity 1) and invariance of
the bound under a permutation of the subjects.
"""
import argparse

import numpy as np
import pytest
import torch

import bench

pytestmark = pytest.mark.gpu

ARGS = argparse.Namespace(path=0, exchange="nccl", no_parity=False)
FULL = [("cfg2", 1000), ("cfg3", 1000), ("cfg4", 2000), ("cfg5", 2000)]


@pytest.mark.parametrize("cfg,spb", FULL)
def test_full_size_parity_vs_oracle_on_device(cfg, spb):
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    r = bench.run_config(ARGS, cfg, spb, 0, 1, dev, None, None, steps=1, warmup=0, parity_only=True)
    p = r.out["parity"]
    assert p["n_subjects_checked_per_rank"] == spb
    assert p["ok"], p
    for k in ("kld", "d_mu", "d_log_v"):                       # outputs that do not amplify Kzz^-1: plain 1e-6
        assert p["per_tensor_rank0"][k] <= 1e-6, p


def _api_bound(cfg, spb, perm=None, use_iter=False):
    import lvae_b200.elbo_functions as EF
    dev = torch.device("cuda", 0)
    b = bench.make_problem(cfg, spb, 0, 1)
    cm0, cm1, lik = bench.build_modules(b, dev)
    T = int(b.T)
    x, mu, lv = b.x, b.mu, b.log_v
    if perm is not None:                      # permute whole subjects
        rows = (np.asarray(perm)[:, None] * T + np.arange(T)[None, :]).reshape(-1)
        x, mu, lv = x[rows], mu[rows], lv[rows]
    c = lambda t: t.to(dev)
    mu_d = c(mu).requires_grad_(True)
    if use_iter:
        kld, gm, gH = EF.minibatch_KLD_upper_bound_iter(cm0, cm1, lik, b.L, c(b.m), c(b.H), c(x), mu_d, c(lv), c(b.z), spb,
                                                        spb, spb * T, True, 2, 1e-6)
    else:
        kld, gm, gH = EF.minibatch_KLD_upper_bound(cm0, cm1, lik, b.L, c(b.m), c(b.H), c(x), mu_d, c(lv), c(b.z), spb, spb, T,
                                                   True, 1e-6)
    kld.sum().backward()
    return float(kld.sum()), gm.detach(), gH.detach(), mu_d.grad.detach()


def test_fixed_T_equals_iter_at_full_size():
    k1, gm1, gH1, dmu1 = _api_bound("cfg2", 1000)
    k2, gm2, gH2, dmu2 = _api_bound("cfg2", 1000, use_iter=True)
    assert abs(k1 - k2) <= 1e-12 * abs(k1)
    assert bench.rel_err(gm2, gm1) < 1e-10 and bench.rel_err(gH2, gH1) < 1e-10 and bench.rel_err(dmu2, dmu1) < 1e-10


def test_subject_permutation_invariance_at_full_size():
    rng = np.random.default_rng(3)
    perm = rng.permutation(1000)
    k1, gm1, gH1, dmu1 = _api_bound("cfg2", 1000)
    k2, gm2, gH2, dmu2 = _api_bound("cfg2", 1000, perm=perm)
    T = 20
    rows = (perm[:, None] * T + np.arange(T)[None, :]).reshape(-1)
    # the order of the subjects only changes the order in which S is summed; Kzz^-1 (cond ~1e8) amplifies that re-ordering
    # noise in grad_m / grad_H, hence the looser bound there
    assert abs(k1 - k2) <= 1e-6 * abs(k1)
    assert bench.rel_err(dmu2, dmu1[torch.from_numpy(rows).to(dmu1.device)]) < 1e-6
    assert bench.rel_err(gH2, gH1) < 1e-4 and bench.rel_err(gm2, gm1) < 1e-4
