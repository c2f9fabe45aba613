"""Shared body of the gradient-parity tests of the non-minibatch bounds: run the drop-in validation_dubo,
deviance_upper_bound, elbo and KL_closed with autograd on `device` and compare values and every gradient with what the
REFERENCE's own functions produced (tests/golden/bounds_grad_*.npz, oracle/make_golden_bounds_grad.py).
Tolerance: 1e-6 relative (FP64) on each value and on each gradient array (max-abs error over max-abs reference)."""
import torch

from conftest import load_golden
from helpers import build_modules, constrained_param_grads, rel

TOL = 1e-6
CASES = ["bounds_grad_cfg2", "bounds_grad_m72", "bounds_grad_cfg4"]


def _unbatched_modules(g, device, l=0):
    from lvae_b200.constraints import GreaterThan
    from lvae_b200.gp_kernels import RBFKernel
    from lvae_b200.kernel_gen import generate_kernel_approx
    from lvae_b200.likelihoods import GaussianLikelihood
    u0, u1 = generate_kernel_approx(**g["lists"], id_covariate=2)
    u0, u1 = u0.double().to(device), u1.double().to(device)
    i_c = i_l = 0
    for mod in (u0, u1):
        for sk in mod.kernels:
            sk.outputscale = torch.as_tensor(g["outputscale"][i_c, l])
            i_c += 1
            for rb in [mm for mm in sk.modules() if isinstance(mm, RBFKernel)]:
                rb.lengthscale = torch.as_tensor(g["lengthscale"][i_l, l])
                i_l += 1
    lik_u = GaussianLikelihood(noise_constraint=GreaterThan(1e-8)).double().to(device)
    lik_u.noise = torch.as_tensor(g["noise"][l])
    return u0, u1, lik_u


def _zero(*mods):
    for m in mods:
        for p in m.parameters():
            p.grad = None


def check_case(name, device):
    import lvae_b200.elbo_functions as EF
    from lvae_b200.validation import validation_dubo
    g = load_golden(name)
    L = g["mu"].shape[1]
    t = lambda k: torch.from_numpy(g[k]).to(device)
    P, T, eps = len(g["offsets"]) - 1, int(g["T"]), float(g["eps"])

    cm0, cm1, lik = build_modules(g["lists"], L, g["lengthscale"], g["outputscale"], g["noise"], device)
    mu, lv = t("mu").requires_grad_(True), t("log_v").requires_grad_(True)
    d = validation_dubo(L, cm0, cm1, lik, t("x"), mu, lv, t("z"), P, T, eps)
    d.sum().backward()
    assert rel(d, g["vdubo"]) < TOL
    assert rel(mu.grad, g["vdubo_d_mu"]) < TOL and rel(lv.grad, g["vdubo_d_log_v"]) < TOL
    assert rel(constrained_param_grads(cm0, cm1, lik), g["vdubo_d_hyper"]) < TOL

    u0, u1, lik_u = _unbatched_modules(g, device)
    mu0 = t("mu")[:, 0].contiguous().requires_grad_(True)
    lv0 = t("log_v")[:, 0].contiguous().requires_grad_(True)
    z0 = t("z")[0].contiguous()
    v = EF.deviance_upper_bound(u0, u1, lik_u, t("x"), mu0, lv0, z0, P, T, eps)
    v.backward()
    assert rel(v, g["dubo0"]) < TOL
    assert rel(mu0.grad, g["dubo0_d_mu"]) < TOL and rel(lv0.grad, g["dubo0_d_log_v"]) < TOL
    assert rel(constrained_param_grads(u0, u1, lik_u), g["dubo0_d_hyper"]) < TOL
    _zero(u0, u1, lik_u)

    y0 = t("mu")[:, 0].contiguous().requires_grad_(True)
    e = EF.elbo(u0, u1, lik_u, t("x"), y0, z0, P, T, eps)
    e.backward()
    assert rel(e, g["elbo0"]) < TOL
    assert rel(y0.grad, g["elbo0_d_y"]) < TOL
    assert rel(constrained_param_grads(u0, u1, lik_u), g["elbo0_d_hyper"]) < TOL
    _zero(u0, u1, lik_u)

    if "klc0" in g:
        mu0 = t("mu")[:, 0].contiguous().requires_grad_(True)
        lv0 = t("log_v")[:, 0].contiguous().requires_grad_(True)
        k = EF.KL_closed(u0 + u1, t("x"), lik_u, mu0, mu0, lv0)
        k.backward()
        assert rel(k, g["klc0"]) < TOL
        assert rel(mu0.grad, g["klc0_d_mu"]) < TOL and rel(lv0.grad, g["klc0_d_log_v"]) < TOL
        assert rel(constrained_param_grads(u0, u1, lik_u), g["klc0_d_hyper"]) < TOL


def check_batched_over_latent_lists(name, device):
    """deviance_upper_bound_all / elbo_all: L un-batched (covar_module0[i], covar_module1[i], likelihoods[i]) triples evaluated
    in one batched call reproduce the reference's per-latent loop — value and gradients of the summed DUBO against
    validation_dubo's golden (same numbers: the batched modules carry the same hyper-parameters), entry 0 of the ELBO vector
    against the single-latent golden."""
    import numpy as np
    import lvae_b200.elbo_functions as EF
    g = load_golden(name)
    L = g["mu"].shape[1]
    t = lambda k: torch.from_numpy(g[k].copy()).to(device)
    P, T, eps = len(g["offsets"]) - 1, int(g["T"]), float(g["eps"])
    trip = [_unbatched_modules(g, device, l) for l in range(L)]
    c0, c1, lk = [a for a, _, _ in trip], [b for _, b, _ in trip], [c for _, _, c in trip]
    zl = [t("z")[l].contiguous() for l in range(L)]
    mu, lv = t("mu").requires_grad_(True), t("log_v").requires_grad_(True)
    v = EF.deviance_upper_bound_all(c0, c1, lk, t("x"), mu, lv, zl, P, T, eps)
    assert v.shape == (L,) and rel(v[0], g["dubo0"]) < TOL and rel(v.sum(), g["vdubo"].sum()) < TOL
    v.sum().backward()
    assert rel(mu.grad, g["vdubo_d_mu"]) < TOL and rel(lv.grad, g["vdubo_d_log_v"]) < TOL
    per_latent = np.stack([constrained_param_grads(*tr) for tr in trip])          # [L, n]; golden blocks are [n][L]
    assert rel(per_latent.T.reshape(-1), g["vdubo_d_hyper"]) < TOL
    for tr in trip:
        _zero(*tr)
    Z = t("mu").requires_grad_(True)
    e = EF.elbo_all(c0, c1, lk, t("x"), Z, zl, P, T, eps)
    assert e.shape == (L,) and rel(e[0], g["elbo0"]) < TOL
    e[0].backward()
    assert rel(Z.grad[:, 0], g["elbo0_d_y"]) < TOL and float(Z.grad[:, 1:].abs().max()) == 0.0
    assert rel(constrained_param_grads(*trip[0]), g["elbo0_d_hyper"]) < TOL
    assert float(np.abs(constrained_param_grads(*trip[1])).max()) == 0.0          # latents are independent
