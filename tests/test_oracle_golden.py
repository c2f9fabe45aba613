"""CPU: the oracle restatement (oracle/lvae_oracle.py) against vectors produced by the reference's own functions
(oracle/make_golden.py).  Tolerances: 1e-9 relative on kld and kernels (same algorithm, FP64, different association),
1e-6 relative-to-max on gradients (north_star FP64 gate)."""
import numpy as np
import torch

import lvae_oracle as orc
from conftest import load_golden, oracle_components


WELL_CONDITIONED = ("cfg2_noNG", "cfg4_ragged", "missing_mask")      # cond(Kzz + eps I) < 1e4


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def run_oracle(g, k0, k1):
    t = lambda k: torch.from_numpy(g[k])
    L = g["mu"].shape[1]
    ng = bool(g["natural_gradient"])
    mu = t("mu").clone().requires_grad_(True)
    lv = t("log_v").clone().requires_grad_(True)
    m = t("m").clone().requires_grad_(not ng)
    H = t("H").clone().requires_grad_(not ng)
    noise = t("noise").clone().requires_grad_(True)
    P_b = len(g["offsets"]) - 1
    if bool(g["ragged"]):
        out = orc.kld_iter(k0, k1, noise, L, m, H, t("x"), mu, lv, t("z"), int(g["P_tot"]), P_b, int(g["N_tot"]), ng, 2,
                           float(g["eps"]))
    else:
        out = orc.kld_fixed_T(k0, k1, noise, L, m, H, t("x"), mu, lv, t("z"), int(g["P_tot"]), P_b, int(g["T"]), ng,
                              float(g["eps"]))
    out[0].backward()
    return out, dict(mu=mu, lv=lv, m=m, H=H, noise=noise)


def test_kernels_match_reference(golden):
    name, g = golden
    k0, k1 = oracle_components(g)
    L = g["mu"].shape[1]
    x, z = torch.from_numpy(g["x"]), torch.from_numpy(g["z"])
    with torch.no_grad():
        assert rel(orc.dense(k0, x, z, L), g["K0xz"]) < 1e-12
        assert rel(orc.dense(k0, z, z, L), g["K0zz"]) < 1e-12
        assert rel(orc.dense(k0, x[:40].unsqueeze(0).expand(L, -1, -1), z, L), g["GPmodel_K0xz"]) < 1e-12
        for p in range(2):
            xs = x[g["offsets"][p]:g["offsets"][p + 1]].unsqueeze(0).expand(L, -1, -1)
            assert rel(orc.dense(k0, xs, xs, L), g[f"K0_block{p}"]) < 1e-12
            assert rel(orc.dense(k1, xs, xs, L), g[f"K1_block{p}"]) < 1e-12
    # categorical / binary structure is exact: masks are 0/1 combinations scaled by outputscale -> compare zero pattern
    assert np.array_equal(orc.dense(k1, xs, xs, L).detach().numpy() == 0, g["K1_block1"] == 0)


def test_bound_and_gradients_match_reference(golden):
    name, g = golden
    k0, k1 = oracle_components(g)
    (kld, gm, gH), leaves = run_oracle(g, k0, k1)
    assert abs(kld.item() - float(g["kld"])) <= 1e-9 * abs(float(g["kld"]))
    assert rel(leaves["mu"].grad, g["d_mu"]) < 1e-7
    assert rel(leaves["lv"].grad, g["d_log_v"]) < 1e-7
    if bool(g["natural_gradient"]):
        assert rel(gm.detach(), g["grad_m"]) < 1e-7
        assert rel(gH.detach(), g["grad_H"]) < 1e-7
    else:
        assert rel(leaves["m"].grad, g["d_m"]) < 1e-7
        assert rel(leaves["H"].grad, g["d_H"]) < 1e-7
    # Hyper-parameter gradients: norm-wise 1e-6 over the whole kernel-parameter gradient of the step.  Entry-wise the
    # small K1 gradients of ill-conditioned cases (cond(Kzz) >= 5e6: cfg1_small, cfg3_small) differ by 1e-6..5e-4 even
    # between two FP64 evaluations of the reference's own formulas (direct-difference SE here vs GPyTorch's
    # norm-expansion SE in the golden run, a 1e-15 perturbation of Kzz) — the reference's arithmetic is only
    # conditionally stable there, so entry-wise 1e-9 is asserted on the well-conditioned cases only.
    mine = np.concatenate([p.grad.numpy().ravel() for comp in k0 + k1 for p in comp.params()]
                          + [leaves["noise"].grad.numpy().ravel()])
    n_par = sum(len(comp.params()) for comp in k0 + k1)
    ref = np.concatenate([g[f"d_param_{i}"].ravel() for i in range(n_par)] + [g["d_noise"].ravel()])
    assert np.abs(mine - ref).max() <= 1e-6 * np.abs(ref).max(), name
    if name in WELL_CONDITIONED:
        assert np.all(np.abs(mine - ref) <= 1e-9 * np.abs(ref) + 1e-12), name


def test_fixed_T_equals_iter_on_regular_input():
    """SURVEY 4 identity 1."""
    g = load_golden("cfg2_small")
    k0, k1 = oracle_components(g)
    t = lambda k: torch.from_numpy(g[k])
    L, P_b, T = g["mu"].shape[1], len(g["offsets"]) - 1, int(g["T"])
    with torch.no_grad():
        a = orc.kld_fixed_T(k0, k1, t("noise"), L, t("m"), t("H"), t("x"), t("mu"), t("log_v"), t("z"), 15, P_b, T, True, 1e-6)
        b = orc.kld_iter(k0, k1, t("noise"), L, t("m"), t("H"), t("x"), t("mu"), t("log_v"), t("z"), 15, P_b, 15 * T, True, 2, 1e-6)
    assert abs(a[0].item() - b[0].item()) <= 1e-9 * abs(a[0].item())
    assert rel(a[1], b[1]) < 1e-9 and rel(a[2], b[2]) < 1e-9


def test_natural_gradient_fixed_point():
    """SURVEY 4 identity 3: full batch, one NG step with lr=1 => next grad_m, grad_H vanish."""
    g = load_golden("cfg2_small")
    k0, k1 = oracle_components(g)
    t = lambda k: torch.from_numpy(g[k])
    L, P_b, T = g["mu"].shape[1], len(g["offsets"]) - 1, int(g["T"])
    args = (t("x"), t("mu"), t("log_v"), t("z"), P_b, P_b, T, True, 1e-6)
    with torch.no_grad():
        _, gm, gH = orc.kld_fixed_T(k0, k1, t("noise"), L, t("m"), t("H"), *args)
        m1, H1 = orc.ng_step(t("m"), t("H"), gm, gH, 1.0)
        _, gm1, gH1 = orc.kld_fixed_T(k0, k1, t("noise"), L, m1, H1, *args)
    assert gm1.abs().max() < 1e-6 * gm.abs().max() and gH1.abs().max() < 1e-6 * gH.abs().max()


def test_samplers_match_reference():
    s = dict(np.load("tests/golden/samplers.npz"))
    for k in range(3):
        P, T, spb = s[f"fixed{k}_PTspb"]
        assert orc.subject_sampler_rows(s[f"fixed{k}_perm"], T) == s[f"fixed{k}_rows"].tolist()
        bl = [len(b) for b in orc.fixed_T_batches(s[f"fixed{k}_perm"], T, spb)]
        assert bl == s[f"fixed{k}_batch_lens"].tolist()
    for k in range(2):
        batches = orc.varying_T_batches(s["vary_ids"], s[f"vary{k}_perm"], int(s[f"vary{k}_spb"]))
        assert [i for b in batches for i in b] == s[f"vary{k}_flat"].tolist()
        assert [len(b) for b in batches] == s[f"vary{k}_batch_lens"].tolist()
