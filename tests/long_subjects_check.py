"""Shared body: minibatch_KLD_upper_bound_iter / minibatch_KLD_upper_bound on subjects with 41..55 rows against the oracle."""
import numpy as np
import torch

from helpers import build_modules, constrained_param_grads, rel

TOL = 1e-6


def check_long_subjects(device, ng):
    import lvae_oracle as orc
    import lvae_b200.elbo_functions as EF
    from lvae_b200 import synth
    P, L, M = 5, 2, 12
    b = synth.make_batch("cfg4", P=P, L=L, M=M, T=(38, 55), seed=21)
    assert int(np.diff(b.offsets).max()) > 40
    z = b.z.clone()
    z[:, :, 0] += 0.21 * torch.arange(M, dtype=torch.float64)                   # distinct inducing inputs
    k0, k1 = orc.parse_kernel_lists(L, **b.lists, id_covariate=2)
    n_ls = sum(len(c.lengthscales) for c in k0 + k1)
    ls, os_, noise = synth.perturbed_hypers(n_ls, len(k0) + len(k1), L, seed=9, noise_trainable=True)
    leaves = []
    i_ls = 0
    for i_c, comp in enumerate(k0 + k1):
        comp.outputscale = os_[i_c].clone().requires_grad_(True)
        leaves.append(comp.outputscale)
        for k in sorted(comp.lengthscales):
            comp.lengthscales[k] = ls[i_ls].clone().requires_grad_(True)
            leaves.append(comp.lengthscales[k])
            i_ls += 1
    noise_o = noise.clone().requires_grad_(True)
    mu_o, lv_o = b.mu.clone().requires_grad_(True), b.log_v.clone().requires_grad_(True)
    m_o, H_o = b.m.clone().requires_grad_(not ng), b.H.clone().requires_grad_(not ng)
    ref = orc.kld_iter(k0, k1, noise_o, L, m_o, H_o, b.x, mu_o, lv_o, z, 3 * P, P, 3 * b.N, ng, 2, 1e-6)
    ref[0].sum().backward()
    ref_hyper = torch.cat([t.grad.reshape(-1) for t in leaves] + [noise_o.grad.reshape(-1)]).numpy()

    cm0, cm1, lik = build_modules(b.lists, L, ls.numpy(), os_.numpy(), noise.numpy(), device)
    c = lambda t: t.clone().to(device)
    mu, lv = c(b.mu).requires_grad_(True), c(b.log_v).requires_grad_(True)
    m, H = c(b.m).requires_grad_(not ng), c(b.H).requires_grad_(not ng)
    perm = torch.randperm(b.N, generator=torch.Generator().manual_seed(2))      # rows of a subject need not be contiguous
    xp = c(b.x)[perm.to(device)]
    mu_p, lv_p = mu[perm.to(device)], lv[perm.to(device)]
    out = EF.minibatch_KLD_upper_bound_iter(cm0, cm1, lik, L, m, H, xp, mu_p, lv_p, c(z), 3 * P, P, 3 * b.N, ng, 2, 1e-6)
    out[0].sum().backward()
    assert abs(float(out[0].detach().sum()) - float(ref[0].detach().sum())) <= TOL * abs(float(ref[0].detach().sum()))
    assert rel(mu.grad, mu_o.grad) < TOL and rel(lv.grad, lv_o.grad) < TOL
    assert rel(constrained_param_grads(cm0, cm1, lik), ref_hyper) < TOL
    if ng:
        assert rel(out[1], ref[1]) < TOL and rel(out[2], ref[2]) < TOL
    else:
        assert out[1] is None and out[2] is None
        assert rel(m.grad, m_o.grad) < TOL and rel(H.grad, H_o.grad) < TOL
    # fixed T above the limit through the fixed-T entry point (constant term L * P_tot * T / 2, elbo_functions.py:204)
    T = 44
    bf = synth.make_batch("cfg2", P=3, L=L, M=M, T=T, seed=5)
    k0f, k1f = orc.parse_kernel_lists(L, **bf.lists, id_covariate=2)
    n_lsf = sum(len(cc.lengthscales) for cc in k0f + k1f)
    lsf, osf, nf = synth.perturbed_hypers(n_lsf, len(k0f) + len(k1f), L, seed=3)
    i_ls = 0
    for i_c, comp in enumerate(k0f + k1f):
        comp.outputscale = osf[i_c].clone()
        for k in sorted(comp.lengthscales):
            comp.lengthscales[k] = lsf[i_ls].clone()
            i_ls += 1
    zf = bf.z.clone()
    zf[:, :, 0] += 0.21 * torch.arange(M, dtype=torch.float64)
    reff = orc.kld_fixed_T(k0f, k1f, nf, L, bf.m, bf.H, bf.x, bf.mu, bf.log_v, zf, 9, 3, T, True, 1e-6)
    cm0f, cm1f, likf = build_modules(bf.lists, L, lsf.numpy(), osf.numpy(), nf.numpy(), device)
    with torch.no_grad():
        outf = EF.minibatch_KLD_upper_bound(cm0f, cm1f, likf, L, c(bf.m), c(bf.H), c(bf.x), c(bf.mu), c(bf.log_v), c(zf), 9, 3,
                                            T, True, 1e-6)
    assert abs(float(outf[0].sum()) - float(reff[0].sum())) <= TOL * abs(float(reff[0].sum()))
    assert rel(outf[1], reff[1]) < TOL and rel(outf[2], reff[2]) < TOL
