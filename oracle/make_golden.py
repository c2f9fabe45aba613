"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the REFERENCE's own functions on CPU FP64.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

Imports, unmodified, `/root/reference/{elbo_functions,kernel_spec,kernel_gen,GP_model,utils}.py`; `gpytorch` resolves to
the stand-in in oracle/gpytorch_standin.  For every case it stores the inputs, the reference outputs
(kernel matrices, kld_total, grad_m, grad_H), autograd gradients w.r.t. mu, log_v and the constrained
hyper-parameters, and sampler index lists.  The committed vectors pin oracle/lvae_oracle.py (tests/test_oracle_golden.py)
and, on the GPU, the CUDA path (tests/test_gpu_parity.py).
"""
import importlib.util
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LVAE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "gpytorch_standin"))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    import gpytorch  # noqa: F401  (the stand-in)
    ref = {}
    for name in ("kernel_spec", "kernel_gen", "elbo_functions", "GP_model", "utils"):
        ref[name] = _load(name, os.path.join(REF, name + ".py"))
    return ref


def _load_synth():
    return _load("lvae_synth", os.path.join(ROOT, "longitudinal-vae_b200", "synth.py"))


def set_ref_params(module, comps):
    """Copy constrained (outputscale, lengthscale) values of oracle components into a reference AdditiveKernel."""
    import gpytorch
    assert len(module.kernels) == len(comps)
    for sk, comp in zip(module.kernels, comps):
        sk.outputscale = comp.outputscale.detach().clone()
        rbfs = [mm for mm in sk.modules() if isinstance(mm, gpytorch.kernels.RBFKernel)]
        keys = sorted(comp.lengthscales)
        assert len(rbfs) == len(keys)
        for rb, k in zip(rbfs, keys):
            rb.lengthscale = comp.lengthscales[k].detach().clone().view(-1, 1, 1)


def ref_param_grads(module, comps):
    """d/d(constrained) from the reference's raw-parameter .grad: raw -> softplus -> value, so d/dvalue = d/draw / sigmoid(raw)."""
    import gpytorch
    out = []
    for sk, comp in zip(module.kernels, comps):
        g = sk.raw_outputscale.grad / torch.sigmoid(sk.raw_outputscale.detach())
        out.append(g.reshape(-1).clone())
        rbfs = [mm for mm in sk.modules() if isinstance(mm, gpytorch.kernels.RBFKernel)]
        for rb in rbfs:
            g = rb.raw_lengthscale.grad / torch.sigmoid(rb.raw_lengthscale.detach())
            out.append(g.reshape(-1).clone())
    return out


def main():
    import lvae_oracle as orc
    ref = load_reference()
    synth = _load_synth()
    import gpytorch
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_default_dtype(torch.float32)

    missing_lists = dict(cat_kernel=[2], bin_kernel=[5], sqexp_kernel=[0, 1],
                         cat_int_kernel=[{'cont_covariate': 1, 'cat_covariate': 2},
                                         {'cont_covariate': 0, 'cat_covariate': 3}],
                         bin_int_kernel=[{'cont_covariate': 1, 'bin_covariate': 5}],
                         covariate_missing_val=[{'covariate': 1, 'mask': 4}])
    cases = [
        # name, cfg, overrides, ragged?, natural_gradient, trainable noise, lists override
        ("cfg1_small", "cfg1", dict(P=6, L=3, M=12), False, True, False, None),
        ("cfg2_small", "cfg2", dict(P=5, L=4, M=16), False, True, True, None),
        ("cfg2_noNG", "cfg2", dict(P=4, L=2, M=9), False, False, True, None),
        ("cfg4_ragged", "cfg4", dict(P=7, L=3, M=10), True, True, True, None),
        ("missing_mask", "cfg2", dict(P=5, L=2, M=11), False, True, True, missing_lists),
        ("cfg3_small", "cfg3", dict(P=4, L=2, M=72), False, True, False, None),
    ]
    for name, cfg, ov, ragged, ng, noise_tr, lists in cases:
        b = synth.make_batch(cfg, **ov)
        if lists is not None:
            b.lists = lists
        L, M = b.L, b.M
        k0, k1 = orc.parse_kernel_lists(L, **b.lists, id_covariate=synth.ID_COVARIATE)
        n_ls = sum(len(c.lengthscales) for c in k0 + k1)
        ls, os_, noise = synth.perturbed_hypers(n_ls, len(k0) + len(k1), L, seed=1234, noise_trainable=noise_tr)
        i_ls = 0
        for i_c, comp in enumerate(k0 + k1):
            comp.outputscale = os_[i_c].clone()
            for k in sorted(comp.lengthscales):
                comp.lengthscales[k] = ls[i_ls].clone()
                i_ls += 1
        # reference modules (gpytorch-style, kernel_gen.py) with the same constrained values
        cm0, cm1 = ref["kernel_gen"].generate_kernel_batched(L, **b.lists, id_covariate=synth.ID_COVARIATE)
        cm0.double(), cm1.double()
        set_ref_params(cm0, k0)
        set_ref_params(cm1, k1)
        lik = gpytorch.likelihoods.GaussianLikelihood(batch_shape=torch.Size([L]),
                                                      noise_constraint=gpytorch.constraints.GreaterThan(1e-8)).double()
        lik.noise = noise.view(L, 1)
        mu = b.mu.clone().requires_grad_(True)
        lv = b.log_v.clone().requires_grad_(True)
        m = b.m.clone().requires_grad_(not ng)
        H = b.H.clone().requires_grad_(not ng)
        eps = 1e-6
        P_tot = b.P * 3          # pretend the batch is a third of the data set, so P_tot/P_b != 1
        if ragged:
            N_tot = b.N * 3
            kld, gm, gH = ref["elbo_functions"].minibatch_KLD_upper_bound_iter(
                cm0, cm1, lik, L, m, H, b.x, mu, lv, b.z, P_tot, b.P, N_tot, ng, synth.ID_COVARIATE, eps)
        else:
            kld, gm, gH = ref["elbo_functions"].minibatch_KLD_upper_bound(
                cm0, cm1, lik, L, m, H, b.x, mu, lv, b.z, P_tot, b.P, b.T, ng, eps)
            # identity 1 of SURVEY 4: fixed-T == iter on regular input with N = P_tot*T
            kld_it, gm_it, gH_it = ref["elbo_functions"].minibatch_KLD_upper_bound_iter(
                cm0, cm1, lik, L, m, H, b.x, mu, lv, b.z, P_tot, b.P, P_tot * b.T, ng, synth.ID_COVARIATE, eps)
            assert abs(kld_it.item() - kld.item()) <= 1e-9 * abs(kld.item()), (kld_it.item(), kld.item())
        kld.sum().backward()
        d = dict(
            x=b.x.numpy(), offsets=b.offsets, mu=b.mu.numpy(), log_v=b.log_v.numpy(), z=b.z.numpy(), m=b.m.numpy(),
            H=b.H.numpy(), lengthscale=ls.numpy(), outputscale=os_.numpy(), noise=noise.numpy(),
            P_tot=np.int64(P_tot), eps=np.float64(eps), natural_gradient=np.bool_(ng), ragged=np.bool_(ragged),
            N_tot=np.int64(b.N * 3), T=np.int64(b.T if not ragged else -1),
            kld=kld.detach().reshape(()).numpy(), d_mu=mu.grad.numpy(), d_log_v=lv.grad.numpy(),
            K0xz=cm0(b.x, b.z).evaluate().detach().numpy(), K0zz=cm0(b.z, b.z).evaluate().detach().numpy(),
            d_noise=(lik.noise_covar.raw_noise.grad / torch.sigmoid(lik.noise_covar.raw_noise.detach())).reshape(-1).numpy(),
            lists=np.array(repr(b.lists)),
        )
        if ng:
            d["grad_m"] = gm.detach().numpy()
            d["grad_H"] = gH.detach().numpy()
        else:
            d["d_m"] = m.grad.numpy()
            d["d_H"] = H.grad.numpy()
        for i, g in enumerate(ref_param_grads(cm0, k0) + ref_param_grads(cm1, k1)):
            d[f"d_param_{i}"] = g.numpy()      # order: per component [outputscale, lengthscales...], K0 comps then K1
        # per-subject blocks of both kernels (reference layout [L,T,T] per subject), first two subjects
        for p in range(2):
            xs = b.x[b.offsets[p]:b.offsets[p + 1]].unsqueeze(0).expand(L, -1, -1)
            d[f"K0_block{p}"] = cm0(xs, xs).evaluate().detach().numpy()
            d[f"K1_block{p}"] = cm1(xs, xs).evaluate().detach().numpy()
        # the authors' gpytorch-free restatement (GP_model.py) on the same constrained values, [L,n,Q] inputs
        g0, g1 = ref["GP_model"].generate_kernel_batched(L, **b.lists, id_covariate=synth.ID_COVARIATE)
        for mod, comps in ((g0, k0), (g1, k1)):
            for sk, comp in zip(mod.kernels, comps):
                sk.double()
                sk.scale = comp.outputscale.clone()
                rbfs = [mm for mm in sk.modules() if isinstance(mm, ref["GP_model"].RbfKernel)]
                for rb, k in zip(rbfs, sorted(comp.lengthscales)):
                    rb.lengthscale = comp.lengthscales[k].clone()
        xL = b.x[:40].unsqueeze(0).expand(L, -1, -1)
        d["GPmodel_K0xz"] = g0(xL, b.z).detach().numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(f"{name}: kld={kld.item():.12e}  N={b.N} L={L} M={M}")

    # sampler golden vectors (utils.py:40-113) under fixed numpy seeds
    U = ref["utils"]
    samp = {}
    # torch>=2.2 removed Sampler.__init__(data_source); the reference still passes it (utils.py:46,67,95) — shim only
    torch.utils.data.sampler.Sampler.__init__ = lambda self, *a, **k: None
    for seed, (P, T, spb) in enumerate([(7, 4, 3), (10, 20, 4), (5, 3, 5)]):
        np.random.seed(100 + seed)
        rows = list(iter(U.SubjectSampler(list(range(P * T)), P, T)))
        np.random.seed(100 + seed)
        perm = np.arange(P)
        np.random.shuffle(perm)
        samp[f"fixed{seed}_rows"] = np.array(rows)
        samp[f"fixed{seed}_perm"] = perm
        samp[f"fixed{seed}_PTspb"] = np.array([P, T, spb])
        bs = list(torch.utils.data.sampler.BatchSampler(rows, spb * T, drop_last=False))
        samp[f"fixed{seed}_batch_lens"] = np.array([len(v) for v in bs])
    rng = np.random.default_rng(5)
    lens = rng.integers(2, 7, size=9)
    ids = np.repeat(np.array([4, 9, 1, 7, 3, 8, 2, 6, 5]), lens)          # ids in order of appearance, not sorted
    data = [{'label': torch.tensor([0.0, 0.0, float(i)])} for i in ids]
    for seed, spb in enumerate([2, 4]):
        vs = U.VaryingLengthSubjectSampler(data, 2)
        np.random.seed(200 + seed)
        batches = list(iter(U.VaryingLengthBatchSampler(vs, spb)))
        np.random.seed(200 + seed)
        perm = np.arange(vs.P)
        np.random.shuffle(perm)
        samp[f"vary{seed}_perm"] = perm
        samp[f"vary{seed}_spb"] = np.array(spb)
        samp[f"vary{seed}_flat"] = np.array([i for bb in batches for i in bb])
        samp[f"vary{seed}_batch_lens"] = np.array([len(bb) for bb in batches])
    samp["vary_ids"] = ids
    np.savez_compressed(os.path.join(out_dir, "samplers.npz"), **samp)
    print("samplers ok")


if __name__ == "__main__":
    main()
