"""TEST INFRASTRUCTURE ONLY — golden values AND autograd gradients of the non-minibatch bounds: runs the
REFERENCE's own validation.validation_dubo, elbo_functions.deviance_upper_bound / elbo / KL_closed on CPU FP64 over the
gpytorch stand-in, calls .backward() on each, and stores inputs, values and gradients in tests/golden/bounds_grad_*.npz.
Build container only (needs /root/reference):

    python oracle/make_golden_bounds_grad.py

Gradients w.r.t. kernel hyper-parameters and noise are stored per constrained value (d/draw divided by sigmoid(raw), as in
make_golden.ref_param_grads): one vector in the order [per component: outputscale, its lengthscales ...; then noise].
torch.solve was removed from torch 2.11; the reference still calls it (elbo_functions.py:75,129): shimmed here only.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def _hyper_grads(mods_comps, lik):
    out = []
    for mod, comps in mods_comps:
        out += MG.ref_param_grads(mod, comps)
    rn = lik.noise_covar.raw_noise
    out.append((rn.grad / torch.sigmoid(rn.detach())).reshape(-1).clone())
    return torch.cat(out).numpy()


def _zero(mods, lik):
    for m in list(mods) + [lik]:
        for p in m.parameters():
            p.grad = None


def main():
    import lvae_oracle as orc
    ref = MG.load_reference()
    synth = MG._load_synth()
    import gpytorch
    torch.solve = lambda b, A: (torch.linalg.solve(A, b), None)
    V = MG._load("validation", os.path.join(MG.REF, "validation.py"))
    EFr = ref["elbo_functions"]
    out_dir = os.path.join(MG.ROOT, "tests", "golden")
    cases = [("bounds_grad_cfg2", "cfg2", dict(P=6, L=3, M=14)),
             ("bounds_grad_m72", "cfg3", dict(P=5, L=2, M=72)),
             ("bounds_grad_cfg4", "cfg4", dict(P=5, L=2, M=16, T=12))]
    for name, cfg, ov in cases:
        T_fixed = ov.pop("T", None)
        b = synth.make_batch(cfg, **ov) if T_fixed is None else synth.make_batch(cfg, T=T_fixed, **ov)
        L = b.L
        assert len(set(np.diff(b.offsets).tolist())) == 1, "the non-minibatch bounds need one T for all subjects"
        T = int(b.offsets[1] - b.offsets[0])
        k0, k1 = orc.parse_kernel_lists(L, **b.lists, id_covariate=synth.ID_COVARIATE)
        n_ls = sum(len(c.lengthscales) for c in k0 + k1)
        ls, os_, noise = synth.perturbed_hypers(n_ls, len(k0) + len(k1), L, seed=2468, noise_trainable=True)
        i_ls = 0
        for i_c, comp in enumerate(k0 + k1):
            comp.outputscale = os_[i_c].clone()
            for k in sorted(comp.lengthscales):
                comp.lengthscales[k] = ls[i_ls].clone()
                i_ls += 1
        cm0, cm1 = ref["kernel_gen"].generate_kernel_batched(L, **b.lists, id_covariate=synth.ID_COVARIATE)
        cm0.double(), cm1.double()
        MG.set_ref_params(cm0, k0)
        MG.set_ref_params(cm1, k1)
        lik = gpytorch.likelihoods.GaussianLikelihood(batch_shape=torch.Size([L]),
                                                      noise_constraint=gpytorch.constraints.GreaterThan(1e-8)).double()
        lik.noise = noise.view(L, 1)
        eps = 1e-6
        d = {}
        # --- validation_dubo: batched modules, all latents --------------------------------------------------------------
        mu, lv = b.mu.clone().requires_grad_(True), b.log_v.clone().requires_grad_(True)
        val = V.validation_dubo(L, cm0, cm1, lik, b.x, mu, lv, b.z, b.P, T, eps)
        val.sum().backward()
        d.update(vdubo=val.detach().numpy(), vdubo_d_mu=mu.grad.numpy(), vdubo_d_log_v=lv.grad.numpy(),
                 vdubo_d_hyper=_hyper_grads([(cm0, k0), (cm1, k1)], lik))
        # --- un-batched modules carrying latent 0's hyper-parameters ----------------------------------------------------
        u0, u1 = ref["kernel_gen"].generate_kernel_approx(**b.lists, id_covariate=synth.ID_COVARIATE)
        u0.double(), u1.double()
        k0u, k1u = orc.parse_kernel_lists(1, **b.lists, id_covariate=synth.ID_COVARIATE)
        for mod, comps, compsu in ((u0, k0, k0u), (u1, k1, k1u)):
            for sk, comp, cu in zip(mod.kernels, comps, compsu):
                sk.outputscale = comp.outputscale[0].detach().clone()
                rbfs = [mm for mm in sk.modules() if isinstance(mm, gpytorch.kernels.RBFKernel)]
                for rb, k in zip(rbfs, sorted(comp.lengthscales)):
                    rb.lengthscale = comp.lengthscales[k][0].detach().clone()
        lik_u = gpytorch.likelihoods.GaussianLikelihood(noise_constraint=gpytorch.constraints.GreaterThan(1e-8)).double()
        lik_u.noise = noise[0]
        pairs_u = [(u0, k0u), (u1, k1u)]

        mu0, lv0 = b.mu[:, 0].clone().requires_grad_(True), b.log_v[:, 0].clone().requires_grad_(True)
        v = EFr.deviance_upper_bound(u0, u1, lik_u, b.x, mu0, lv0, b.z[0], b.P, T, eps)
        v.sum().backward()
        d.update(dubo0=v.detach().reshape(()).numpy(), dubo0_d_mu=mu0.grad.numpy(), dubo0_d_log_v=lv0.grad.numpy(),
                 dubo0_d_hyper=_hyper_grads(pairs_u, lik_u))
        _zero([u0, u1], lik_u)

        y0 = b.mu[:, 0].clone().requires_grad_(True)
        e = EFr.elbo(u0, u1, lik_u, b.x, y0, b.z[0], b.P, T, eps)
        e.sum().backward()
        d.update(elbo0=e.detach().reshape(()).numpy(), elbo0_d_y=y0.grad.numpy(), elbo0_d_hyper=_hyper_grads(pairs_u, lik_u))
        _zero([u0, u1], lik_u)

        if b.x.shape[0] <= 256:
            mu0, lv0 = b.mu[:, 0].clone().requires_grad_(True), b.log_v[:, 0].clone().requires_grad_(True)
            kc = EFr.KL_closed(u0 + u1, b.x, lik_u, mu0, mu0, lv0)
            kc.sum().backward()
            d.update(klc0=kc.detach().reshape(()).numpy(), klc0_d_mu=mu0.grad.numpy(), klc0_d_log_v=lv0.grad.numpy(),
                     klc0_d_hyper=_hyper_grads(pairs_u, lik_u))
        d.update(x=b.x.numpy(), offsets=b.offsets, mu=b.mu.numpy(), log_v=b.log_v.numpy(), z=b.z.numpy(),
                 lengthscale=ls.numpy(), outputscale=os_.numpy(), noise=noise.numpy(), eps=np.float64(eps), T=np.int64(T),
                 lists=np.array(repr(b.lists)))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(f"{name}: vdubo {float(val.sum()):.9e} dubo0 {float(v):.9e} elbo0 {float(e):.9e} "
              f"|d_hyper| {np.abs(d['vdubo_d_hyper']).max():.3e}")


if __name__ == "__main__":
    main()
