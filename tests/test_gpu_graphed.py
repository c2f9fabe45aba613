"""GPU: GraphedHensmanStep (one CUDA-graph replay per step) against the step-by-step public API —
minibatch_KLD_upper_bound + backward + natural_gradient_step — over several consecutive steps with an optimiser on the
hyper-parameters in between.  Same kernels underneath, so agreement is asserted at 1e-12."""
import pytest
import torch

from helpers import rel

pytestmark = pytest.mark.gpu


def _setup(cfg, P, L, M):
    from lvae_b200 import synth
    from lvae_b200.constraints import GreaterThan
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200.likelihoods import GaussianLikelihood
    b = synth.make_batch(cfg, P=P, L=L, M=M)
    cm0, cm1 = generate_kernel_batched(L, **b.lists, id_covariate=2)
    cm0, cm1 = cm0.double().cuda(), cm1.double().cuda()
    lik = GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=GreaterThan(1e-8)).double().cuda()
    return b, cm0, cm1, lik


@pytest.mark.parametrize("cfg,P,L,M,spb", [("cfg2", 12, 4, 20, 4), ("cfg3", 12, 2, 72, 4)])
def test_graphed_step_equals_the_public_api_over_several_steps(cfg, P, L, M, spb):
    import lvae_b200.elbo_functions as EF
    from lvae_b200.graphed import GraphedHensmanStep
    from lvae_b200.training import natural_gradient_step
    lr_ng, eps = 0.05, 1e-6
    runs = []
    for graphed in (False, True):
        b, cm0, cm1, lik = _setup(cfg, P, L, M)
        T = b.T
        m, H, z = b.m.cuda().contiguous(), b.H.cuda().contiguous(), b.z.cuda()
        params = list(cm0.parameters()) + list(cm1.parameters()) + list(lik.parameters())
        opt = torch.optim.Adam(params, lr=1e-2)
        step = GraphedHensmanStep(cm0, cm1, lik, L, m, H, z, P, spb, T, eps, lr_ng) if graphed else None
        rec = []
        for it in range(3):
            rows = slice(it * spb * T, (it + 1) * spb * T)
            x = b.x[rows].cuda()
            mu = b.mu[rows].cuda().requires_grad_(True)
            lv = b.log_v[rows].cuda().requires_grad_(True)
            opt.zero_grad()
            if graphed:
                kld = step(x, mu, lv)
                (2.0 * kld).backward()
            else:
                kld, gm, gH = EF.minibatch_KLD_upper_bound(cm0, cm1, lik, L, m, H, x, mu, lv, z, P, spb, T, True, eps)
                (2.0 * kld.sum()).backward()
                m, H = natural_gradient_step(m, H, gm, gH, lr_ng)
            opt.step()
            rec.append(dict(kld=kld.detach().sum().clone(), d_mu=mu.grad.clone(), d_lv=lv.grad.clone(), m=m.clone(), H=H.clone(),
                            params=torch.cat([p.detach().reshape(-1) for p in params]).clone()))
        if graphed:
            step.check_errors()
        runs.append(rec)
    for a, g in zip(*runs):
        for k in a:
            assert rel(g[k], a[k]) < 1e-12, k


def test_graphed_step_guards():
    from lvae_b200.graphed import GraphedHensmanStep
    b, cm0, cm1, lik = _setup("cfg2", 6, 2, 10)
    m, H, z = b.m.cuda().contiguous(), b.H.cuda().contiguous(), b.z.cuda()
    step = GraphedHensmanStep(cm0, cm1, lik, 2, m, H, z, 6, 3, b.T)
    x, mu, lv = b.x.cuda(), b.mu.cuda().requires_grad_(True), b.log_v.cuda().requires_grad_(True)
    with pytest.raises(RuntimeError, match="rows"):
        step(x, mu, lv)                                              # 6 subjects, built for 3
    n = 3 * b.T
    k1 = step(x[:n], mu[:n], lv[:n])
    step(x[n:], mu[n:], lv[n:])
    with pytest.raises(RuntimeError, match="before backward"):
        k1.backward()                                                # static buffers were overwritten by the second call
    # a non-PD H is reported by the next call (deferred flags), not silently ignored
    H.copy_(-torch.eye(H.shape[-1], dtype=torch.float64, device="cuda").expand_as(H))
    step(x[:n], mu[:n], lv[:n])
    with pytest.raises(RuntimeError, match="positive-definite"):
        step.check_errors()
