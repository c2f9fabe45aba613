"""GPU parity of the prediction path against golden vectors produced by the reference's own
utils.batch_predict / batch_predict_varying_T (oracle/make_golden_predict.py).  Tolerance 1e-6 relative (FP64)."""
import pytest
import torch

from conftest import load_golden
from helpers import build_modules, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["predict_fixed", "predict_ragged", "predict_m72"])
def test_batch_predict_vs_reference_golden(name):
    from lvae_b200 import utils as U
    g = load_golden(name)
    L = g["mu"].shape[1]
    cm0, cm1, lik = build_modules(g["lists"], L, g["lengthscale"], g["outputscale"], g["noise"])
    t = lambda k: torch.from_numpy(g[k]).cuda()
    P = len(g["offsets"]) - 1
    Zv = U.batch_predict_varying_T(L, cm0, cm1, lik, t("x"), t("test_x"), t("mu"), t("z"), 2, float(g["eps"]))
    assert Zv.shape == g["Z_pred"].shape and rel(Zv, g["Z_pred"]) < 1e-6
    if not bool(g["ragged"]):
        Z = U.batch_predict(L, cm0, cm1, lik, t("x"), t("test_x"), t("mu"), t("z"), P, int(g["T"]), 2, float(g["eps"]))
        assert rel(Z, g["Z_pred"]) < 1e-6
    # rows of a subject need not be contiguous for the varying-T variant (boolean-mask grouping, utils.py:160-163)
    perm = torch.randperm(g["x"].shape[0], generator=torch.Generator().manual_seed(1)).cuda()
    Zp = U.batch_predict_varying_T(L, cm0, cm1, lik, t("x")[perm], t("test_x"), t("mu")[perm], t("z"), 2, float(g["eps"]))
    assert rel(Zp, g["Z_pred"]) < 1e-6


def test_dubo_elbo_klclosed_vs_reference_golden():
    """validation_dubo (batched), deviance_upper_bound / elbo / KL_closed (un-batched, latent 0) — forward values against the
    reference's own functions."""
    import lvae_b200.elbo_functions as EF
    from lvae_b200.constraints import GreaterThan
    from lvae_b200.kernel_gen import generate_kernel_approx
    from lvae_b200.likelihoods import GaussianLikelihood
    from lvae_b200.validation import validation_dubo
    from lvae_b200.gp_kernels import RBFKernel
    g = load_golden("predict_fixed")
    L = g["mu"].shape[1]
    cm0, cm1, lik = build_modules(g["lists"], L, g["lengthscale"], g["outputscale"], g["noise"])
    t = lambda k: torch.from_numpy(g[k]).cuda()
    P, T, eps = len(g["offsets"]) - 1, int(g["T"]), float(g["eps"])
    d = validation_dubo(L, cm0, cm1, lik, t("x"), t("mu"), t("log_v"), t("z"), P, T, eps)
    assert abs(d.item() - float(g["dubo_sum"][0])) <= 1e-6 * abs(float(g["dubo_sum"][0]))
    # un-batched modules with latent 0's hyper-parameters
    u0, u1 = generate_kernel_approx(**g["lists"], id_covariate=2)
    u0, u1 = u0.double().cuda(), u1.double().cuda()
    i_c = i_l = 0
    for mod in (u0, u1):
        for sk in mod.kernels:
            sk.outputscale = torch.as_tensor(g["outputscale"][i_c, 0])
            i_c += 1
            for rb in [mm for mm in sk.modules() if isinstance(mm, RBFKernel)]:
                rb.lengthscale = torch.as_tensor(g["lengthscale"][i_l, 0])
                i_l += 1
    lik_u = GaussianLikelihood(noise_constraint=GreaterThan(1e-8)).double().cuda()
    lik_u.noise = torch.as_tensor(g["noise"][0])
    mu0, lv0, z0 = t("mu")[:, 0].contiguous(), t("log_v")[:, 0].contiguous(), t("z")[0].contiguous()
    v = EF.deviance_upper_bound(u0, u1, lik_u, t("x"), mu0, lv0, z0, P, T, eps)
    assert abs(v.item() - float(g["dubo_latent0"])) <= 1e-6 * abs(float(g["dubo_latent0"]))
    e = EF.elbo(u0, u1, lik_u, t("x"), mu0, z0, P, T, eps)
    assert abs(e.item() - float(g["elbo_latent0"])) <= 1e-6 * abs(float(g["elbo_latent0"]))
    k = EF.KL_closed(u0 + u1, t("x"), lik_u, mu0, mu0, lv0)
    assert abs(k.item() - float(g["klclosed_latent0"])) <= 1e-6 * abs(float(g["klclosed_latent0"]))
