"""Shared helpers of the GPU parity tests: build the drop-in modules for a golden case and run the CUDA op."""
import numpy as np
import torch

import lvae_b200.elbo_functions as EF
from lvae_b200.gp_kernels import RBFKernel
from lvae_b200.kernel_gen import generate_kernel_batched
from lvae_b200.likelihoods import GaussianLikelihood
from lvae_b200.constraints import GreaterThan


def rel(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-300))


def set_params(cm0, cm1, lengthscale, outputscale):
    """Copy constrained values (rows in K0-then-K1 component order) into drop-in kernel modules."""
    L = outputscale.shape[1]
    i_c = i_l = 0
    for mod in (cm0, cm1):
        for sk in mod.kernels:
            sk.outputscale = torch.as_tensor(outputscale[i_c])
            i_c += 1
            for rb in [mm for mm in sk.modules() if isinstance(mm, RBFKernel)]:
                rb.lengthscale = torch.as_tensor(lengthscale[i_l]).view(L, 1, 1)
                i_l += 1


def build_modules(lists, L, lengthscale, outputscale, noise, device="cuda"):
    cm0, cm1 = generate_kernel_batched(L, **lists, id_covariate=2)
    cm0, cm1 = cm0.double().to(device), cm1.double().to(device)
    set_params(cm0, cm1, lengthscale, outputscale)
    lik = GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=GreaterThan(1e-8)).double().to(device)
    lik.noise = torch.as_tensor(noise).view(L, 1)
    return cm0, cm1, lik


def constrained_param_grads(cm0, cm1, lik):
    """d/d(constrained value) per parameter in golden order: per component [outputscale, lengthscale...], then noise."""
    out = []
    for mod in (cm0, cm1):
        for sk in mod.kernels:
            out.append((sk.raw_outputscale.grad / torch.sigmoid(sk.raw_outputscale.detach())).reshape(-1))
            for rb in [mm for mm in sk.modules() if isinstance(mm, RBFKernel)]:
                out.append((rb.raw_lengthscale.grad / torch.sigmoid(rb.raw_lengthscale.detach())).reshape(-1))
    rn = lik.noise_covar.raw_noise
    out.append((rn.grad / torch.sigmoid(rn.detach())).reshape(-1))
    return torch.cat(out).cpu().numpy()


def run_cuda_case(g, path=0, device="cuda"):
    """Run the drop-in op on a golden case; returns dict of outputs (numpy) in golden naming."""
    t = lambda k: torch.from_numpy(g[k]).to(device)
    L = g["mu"].shape[1]
    ng = bool(g["natural_gradient"])
    cm0, cm1, lik = build_modules(g["lists"], L, g["lengthscale"], g["outputscale"], g["noise"], device)
    mu, lv = t("mu").requires_grad_(True), t("log_v").requires_grad_(True)
    m, H = t("m").requires_grad_(not ng), t("H").requires_grad_(not ng)
    P_b = len(g["offsets"]) - 1
    EF.set_kernel_path(path)
    try:
        if bool(g["ragged"]):
            kld, gm, gH = EF.minibatch_KLD_upper_bound_iter(cm0, cm1, lik, L, m, H, t("x"), mu, lv, t("z"), int(g["P_tot"]),
                                                            P_b, int(g["N_tot"]), ng, 2, float(g["eps"]))
        else:
            kld, gm, gH = EF.minibatch_KLD_upper_bound(cm0, cm1, lik, L, m, H, t("x"), mu, lv, t("z"), int(g["P_tot"]),
                                                       P_b, int(g["T"]), ng, float(g["eps"]))
    finally:
        EF.set_kernel_path(0)
    kld.sum().backward()
    out = dict(kld=float(kld.sum().item()), d_mu=mu.grad.cpu().numpy(), d_log_v=lv.grad.cpu().numpy(),
               d_hyper=constrained_param_grads(cm0, cm1, lik))
    if ng:
        out["grad_m"], out["grad_H"] = gm.detach().cpu().numpy(), gH.detach().cpu().numpy()
    else:
        out["d_m"], out["d_H"] = m.grad.cpu().numpy(), H.grad.cpu().numpy()
    return out


def golden_hyper_vector(g):
    n = sum(1 for k in g if k.startswith("d_param_"))
    return np.concatenate([g[f"d_param_{i}"].ravel() for i in range(n)] + [g["d_noise"].ravel()])
