"""TEST INFRASTRUCTURE ONLY — golden vectors for the prediction path: runs the REFERENCE's own
utils.batch_predict / utils.batch_predict_varying_T (utils.py:115-299) on CPU FP64 over the gpytorch stand-in and stores
inputs and Z_pred in tests/golden/predict_*.npz.  Build container only (needs /root/reference):

    python oracle/make_golden_predict.py

torch.solve was removed from torch 2.11; the reference still calls it (utils.py:176,185,272,276): shimmed here only.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main():
    import lvae_oracle as orc
    ref = MG.load_reference()
    synth = MG._load_synth()
    import gpytorch
    torch.solve = lambda b, A: (torch.linalg.solve(A, b), None)
    U = ref["utils"]
    out_dir = os.path.join(MG.ROOT, "tests", "golden")
    cases = [("predict_fixed", "cfg2", dict(P=6, L=3, M=14), False),
             ("predict_ragged", "cfg4", dict(P=7, L=2, M=12), True),
             ("predict_m72", "cfg3", dict(P=5, L=2, M=72), False)]
    for name, cfg, ov, ragged in cases:
        b = synth.make_batch(cfg, **ov)
        L = b.L
        k0, k1 = orc.parse_kernel_lists(L, **b.lists, id_covariate=synth.ID_COVARIATE)
        n_ls = sum(len(c.lengthscales) for c in k0 + k1)
        ls, os_, noise = synth.perturbed_hypers(n_ls, len(k0) + len(k1), L, seed=4321, noise_trainable=True)
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
        # test rows: later time points of the first subjects (seen ids) plus two unseen subjects
        rng = np.random.default_rng(7)
        bt = synth.make_batch(cfg, P=b.P + 2, L=L, M=ov["M"], seed=777)
        seen = bt.x[: int(bt.offsets[3])].clone()
        seen[:, 0] += 0.5                                   # shifted ages, same subject ids 0..2
        unseen = bt.x[int(bt.offsets[b.P]):].clone()        # ids P, P+1 do not occur in the prediction set
        test_x = torch.cat([seen, unseen])[torch.from_numpy(rng.permutation(seen.shape[0] + unseen.shape[0]))]
        eps = 1e-6
        with torch.no_grad():
            if ragged:
                Z = U.batch_predict_varying_T(L, cm0, cm1, lik, b.x, test_x, b.mu, b.z, synth.ID_COVARIATE, eps)
            else:
                Z = U.batch_predict(L, cm0, cm1, lik, b.x, test_x, b.mu, b.z, b.P, b.T, synth.ID_COVARIATE, eps)
                Zv = U.batch_predict_varying_T(L, cm0, cm1, lik, b.x, test_x, b.mu, b.z, synth.ID_COVARIATE, eps)
                assert float((Z - Zv).abs().max()) <= 1e-8 * float(Z.abs().max())
        extra = {}
        if not ragged:
            # validation_dubo (validation.py:8-68, batched modules) and the single-latent deviance_upper_bound / elbo /
            # KL_closed (elbo_functions.py:8-142) on un-batched reference kernels carrying latent 0's hyper-parameters
            V = MG._load("validation", os.path.join(MG.REF, "validation.py"))
            with torch.no_grad():
                extra["dubo_sum"] = V.validation_dubo(L, cm0, cm1, lik, b.x, b.mu, b.log_v, b.z, b.P, b.T, eps).numpy()
                u0, u1 = ref["kernel_gen"].generate_kernel_approx(**b.lists, id_covariate=synth.ID_COVARIATE)
                u0.double(), u1.double()
                for mod, comps in ((u0, k0), (u1, k1)):
                    for sk, comp in zip(mod.kernels, comps):
                        sk.outputscale = comp.outputscale[0].detach().clone()
                        rbfs = [mm for mm in sk.modules() if isinstance(mm, gpytorch.kernels.RBFKernel)]
                        for rb, k in zip(rbfs, sorted(comp.lengthscales)):
                            rb.lengthscale = comp.lengthscales[k][0].detach().clone()
                lik_u = gpytorch.likelihoods.GaussianLikelihood(noise_constraint=gpytorch.constraints.GreaterThan(1e-8)).double()
                lik_u.noise = noise[0]
                EFr = ref["elbo_functions"]
                extra["dubo_latent0"] = EFr.deviance_upper_bound(u0, u1, lik_u, b.x, b.mu[:, 0], b.log_v[:, 0], b.z[0], b.P, b.T,
                                                                 eps).reshape(()).numpy()
                extra["elbo_latent0"] = EFr.elbo(u0, u1, lik_u, b.x, b.mu[:, 0], b.z[0], b.P, b.T, eps).reshape(()).numpy()
                extra["klclosed_latent0"] = EFr.KL_closed(u0 + u1, b.x, lik_u, b.mu[:, 0], b.mu[:, 0], b.log_v[:, 0]).reshape(()).numpy()
            extra["log_v"] = b.log_v.numpy()
        d = dict(x=b.x.numpy(), offsets=b.offsets, mu=b.mu.numpy(), z=b.z.numpy(), test_x=test_x.numpy(), **extra,
                 lengthscale=ls.numpy(), outputscale=os_.numpy(), noise=noise.numpy(), eps=np.float64(eps),
                 ragged=np.bool_(ragged), T=np.int64(b.T if not ragged else -1), Z_pred=Z.numpy(),
                 lists=np.array(repr(b.lists)))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(f"{name}: Z_pred {tuple(Z.shape)} max|Z|={float(Z.abs().max()):.6e}")


if __name__ == "__main__":
    main()
