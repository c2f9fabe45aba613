"""GPU parity of the differentiable non-minibatch bounds: values and all gradients of validation_dubo,
deviance_upper_bound, elbo and KL_closed through the CUDA ops against the reference's own autograd results
(tests/golden/bounds_grad_*.npz).  Also the raw hyper-parameter adjoint kernels against a torch restatement."""
import numpy as np
import pytest
import torch

from bounds_grad_check import CASES, check_batched_over_latent_lists, check_case
from conftest import load_golden
from helpers import rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_bounds_values_and_gradients_match_reference_golden(name):
    check_case(name, "cuda")


@pytest.mark.parametrize("name", CASES)
def test_batched_evaluation_of_per_latent_module_lists(name):
    check_batched_over_latent_lists(name, "cuda")


@pytest.mark.parametrize("which", ["k0", "k1", "all"])
def test_kernel_adjoint_ops_match_torch_restatement(which):
    """lvae_kernel_dense_bwd_f64 / lvae_kernel_blocks_bwd_f64 (ragged subjects, stacked [P*L] covariates, diagonal term)
    against autograd through the torch stand-in of the forward op."""
    import ops_emulation as emu
    from lvae_b200 import ops, synth
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200.spec import build_structure, flatten
    L = 3
    b = synth.make_batch("cfg4", P=9, L=L, M=10)
    cm0, cm1 = generate_kernel_batched(L, **b.lists, id_covariate=2)
    st, ls, os_ = build_structure(flatten(cm0), flatten(cm1), L)
    ls_, os_p, _ = synth.perturbed_hypers(st.n_ls, st.n_comp, L, seed=5)
    noise = torch.rand(L, dtype=torch.float64) + 0.5
    x, z = b.x, b.z
    gen = torch.Generator().manual_seed(3)
    dev = lambda t: t.cuda()
    # dense, shared covariates
    G = torch.randn(L, x.shape[0], z.shape[1], generator=gen, dtype=torch.float64)
    want = emu.kernel_dense_bwd(st, ls_, os_p, x, z, G, which)
    got = ops.kernel_dense_bwd(st, dev(ls_), dev(os_p), dev(x), dev(z), dev(G), which)
    for a, w in zip(got[:2], want[:2]):
        assert rel(a, w) < 1e-10 or float(w.abs().max()) == 0.0 and float(a.abs().max()) == 0.0
    # dense, stacked [2*L, n, Q] covariates with the diagonal term
    xs = torch.stack([x[:12] + 0.1 * i for i in range(2 * L)])
    G = torch.randn(2 * L, 12, 12, generator=gen, dtype=torch.float64)
    want = emu.kernel_dense_bwd(st, ls_, os_p, xs, xs, G, which, want_diag=True)
    got = ops.kernel_dense_bwd(st, dev(ls_), dev(os_p), dev(xs), dev(xs), dev(G), which, want_diag=True)
    for a, w in zip(got, want):
        assert rel(a, w) < 1e-10 or float(w.abs().max()) == 0.0 and float(a.abs().max()) == 0.0
    # ragged per-subject blocks
    off = torch.from_numpy(np.asarray(b.offsets, dtype=np.int32))
    sum_T2 = int((np.diff(b.offsets) ** 2).sum())
    G = torch.randn(L, sum_T2, generator=gen, dtype=torch.float64)
    want = emu.kernel_blocks_bwd(st, ls_, os_p, x, off, G, which, want_diag=True)
    got = ops.kernel_blocks_bwd(st, dev(ls_), dev(os_p), dev(x), dev(off), dev(G), which, want_diag=True)
    for a, w in zip(got, want):
        assert rel(a, w) < 1e-10 or float(w.abs().max()) == 0.0 and float(a.abs().max()) == 0.0
    # and the forward ops against the same stand-in
    assert rel(ops.kernel_dense(st, dev(ls_), dev(os_p), dev(x), dev(z), which), emu.kernel_dense(st, ls_, os_p, x, z, which)) < 1e-12
    assert rel(ops.kernel_blocks(st, dev(ls_), dev(os_p), dev(x), dev(off), sum_T2, which, diag_add=dev(noise)),
               emu.kernel_blocks(st, ls_, os_p, x, off, sum_T2, which, diag_add=noise)) < 1e-12


def test_evaluate_is_differentiable_like_gpytorch_lazy_kernels():
    """covar_module(x1, x2).evaluate() under a loss: gradients reach raw_outputscale / raw_lengthscale."""
    from helpers import build_modules, constrained_param_grads
    g = load_golden("bounds_grad_cfg2")
    L = g["mu"].shape[1]
    cm0, cm1, lik = build_modules(g["lists"], L, g["lengthscale"], g["outputscale"], g["noise"])
    x, z = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["z"]).cuda()
    Wt = torch.randn(L, x.shape[0], z.shape[1], dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    ((cm0 + cm1)(x, z).evaluate() * Wt).sum().backward()
    import ops_emulation as emu
    from lvae_b200.spec import build_structure, flatten
    st, ls, os_ = build_structure(flatten(cm0), flatten(cm1), L, device="cuda")
    d_ls, d_os, _ = emu.kernel_dense_bwd(st, ls.detach().cpu(), os_.detach().cpu(), x.cpu(), z.cpu(), Wt.cpu(), "all")
    lik.noise_covar.raw_noise.grad = torch.zeros_like(lik.noise_covar.raw_noise)
    got = constrained_param_grads(cm0, cm1, lik)[:-L]
    # golden order: per component [outputscale, its lengthscales...]
    want, i_l = [], 0
    for c in range(st.n_comp):
        want.append(d_os[c])
        if st.table[c][0] >= 0:
            want.append(d_ls[i_l])
            i_l += 1
    assert rel(got, torch.cat(want)) < 1e-10


def _adam_on_dubo(g, device, steps):
    """A few Adam steps on the summed DUBO over (mu, log_v, all kernel hyper-parameters, noise) — the shape of the reference's
    variational_inference_optimization loop (training.py:640-670) without the VAE around it."""
    from helpers import build_modules
    from lvae_b200.validation import validation_dubo
    L = g["mu"].shape[1]
    t = lambda k: torch.from_numpy(g[k].copy()).to(device)          # copy: Adam updates mu / log_v in place
    P, T, eps = len(g["offsets"]) - 1, int(g["T"]), float(g["eps"])
    cm0, cm1, lik = build_modules(g["lists"], L, g["lengthscale"], g["outputscale"], g["noise"], device)
    mu, lv = t("mu").requires_grad_(True), t("log_v").requires_grad_(True)
    params = [mu, lv] + list(cm0.parameters()) + list(cm1.parameters()) + list(lik.parameters())
    opt = torch.optim.Adam(params, lr=1e-2)
    losses = []
    for _ in range(steps):
        opt.zero_grad()
        loss = validation_dubo(L, cm0, cm1, lik, t("x"), mu, lv, t("z"), P, T, eps).sum()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return losses, mu.detach().cpu(), torch.cat([p.detach().reshape(-1).cpu() for p in params[2:]])


def test_adam_on_dubo_follows_the_cpu_trajectory():
    from ops_emulation import emulated_ops
    g = load_golden("bounds_grad_cfg2")
    with emulated_ops():
        want = _adam_on_dubo(g, "cpu", 6)
    got = _adam_on_dubo(g, "cuda", 6)
    assert got[0][-1] < got[0][0]                                   # it descends
    assert rel(np.asarray(got[0]), np.asarray(want[0])) < 1e-6
    assert rel(got[1], want[1]) < 1e-6 and rel(got[2], want[2]) < 1e-6


@pytest.mark.parametrize("ng", [True, False])
def test_long_subjects_take_the_composed_path_and_match_the_oracle(ng):
    """More than 40 rows per subject (41..55 here): elbo_functions._composed_bound on the CUDA ops against the oracle."""
    from long_subjects_check import check_long_subjects
    check_long_subjects("cuda", ng)
