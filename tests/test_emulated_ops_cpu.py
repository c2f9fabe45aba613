"""CPU: autograd plumbing of the differentiable non-minibatch bounds (diff_ops.py + elbo_functions.py) against the
reference's golden gradients, with the C-ABI ops swapped for torch stand-ins (tests/ops_emulation.py).  What this pins
without a GPU: every adjoint formula above the ops (SPD inverse / log-det, the GEMM transposes, which kernel matrix feeds
which term).  The CUDA kernels under the same Functions are checked by tests/test_gpu_bounds_grad.py."""
import pytest
import torch

from bounds_grad_check import CASES, check_batched_over_latent_lists, check_case
from ops_emulation import emulated_ops


@pytest.mark.parametrize("name", CASES)
def test_bounds_values_and_gradients_match_reference_golden(name):
    with emulated_ops():
        check_case(name, "cpu")


@pytest.mark.parametrize("name", CASES)
def test_batched_evaluation_of_per_latent_module_lists(name):
    with emulated_ops():
        check_batched_over_latent_lists(name, "cpu")


def test_emulation_is_removed_afterwards_and_product_path_fails_loudly():
    import lvae_b200.elbo_functions as EF
    from lvae_b200 import ops
    with emulated_ops():
        pass
    assert ops.kernel_dense.__module__ == "lvae_b200.ops"
    x = torch.zeros(4, 6, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="CUDA"):
        EF.deviance_upper_bound(None, None, None, x, x[:, 0], x[:, 0], x[:2], 2, 2, 1e-6)


def test_bounds_gradients_agree_with_central_differences():
    """Independent of any golden file: d(sum DUBO)/d(theta) and d(sum ELBO)/d(theta) from the autograd Functions of diff_ops.py
    against central differences, for every raw kernel / noise parameter and for entries of mu, log_v and the latent sample
    (tiny well-conditioned problem, torch stand-ins for the C-ABI ops)."""
    import lvae_b200.elbo_functions as EF
    from lvae_b200 import synth
    from lvae_b200.constraints import GreaterThan
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200.likelihoods import GaussianLikelihood
    L, P, T, M = 2, 3, 4, 5
    b = synth.make_batch("cfg4", P=P, L=L, M=M, T=T)
    z = b.z.clone()
    z[:, :, 0] += 0.37 * torch.arange(M, dtype=torch.float64)          # distinct inducing inputs: well-conditioned Kzz
    torch.manual_seed(0)
    cm0, cm1 = generate_kernel_batched(L, **b.lists, id_covariate=2)
    cm0, cm1 = cm0.double(), cm1.double()
    lik = GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=GreaterThan(1e-8)).double()
    params = [p for m in (cm0, cm1, lik) for p in m.parameters()]
    with torch.no_grad():
        for p in params:
            p.add_(0.3 * torch.randn_like(p))
    mu, lv = b.mu.clone().requires_grad_(True), b.log_v.clone().requires_grad_(True)

    def dubo():
        return EF._dubo_per_latent(L, cm0, cm1, lik, b.x, mu, lv, z, P, T, 1e-3).sum()

    def elbo():
        return EF._elbo_per_latent(L, cm0, cm1, lik, b.x, mu, z, P, T, 1e-3).sum()

    with emulated_ops():
        for f, leaves in ((dubo, [mu, lv]), (elbo, [mu])):
            for t in params + leaves:
                t.grad = None
            f().backward()
            for t in params + leaves:
                g = t.grad.reshape(-1)
                flat = t.data.reshape(-1)
                for idx in sorted({0, flat.numel() // 2, flat.numel() - 1}):
                    h = 1e-5
                    old = float(flat[idx])
                    with torch.no_grad():
                        flat[idx] = old + h
                        fp = float(f())
                        flat[idx] = old - h
                        fm = float(f())
                        flat[idx] = old
                    fd = (fp - fm) / (2 * h)
                    assert abs(fd - float(g[idx])) <= 1e-6 * max(1.0, abs(fd)), (f.__name__, tuple(t.shape), idx, fd, float(g[idx]))


@pytest.mark.parametrize("ng", [True, False])
def test_long_subjects_take_the_composed_path_and_match_the_oracle(ng):
    """Subjects with more than 40 rows (the fused kernels' limit) go through elbo_functions._composed_bound, built from the
    differentiable ops; with the ops swapped for torch stand-ins its formulas are checked here against the oracle's restatement
    of minibatch_KLD_upper_bound_iter (elbo_functions.py:219-307): bound, natural gradients / d m, d H, d mu, d log_v and
    the hyper-parameter gradients.  The same call on the CUDA ops: tests/test_gpu_bounds_grad.py."""
    from long_subjects_check import check_long_subjects
    with emulated_ops():
        check_long_subjects("cpu", ng)


@pytest.mark.parametrize("name", ["predict_fixed", "predict_ragged", "predict_m72"])
def test_prediction_host_logic_matches_reference_golden(name):
    """utils.batch_predict / batch_predict_varying_T above the ops (grouping by rows per subject, scatter back, the seen /
    unseen-subject split of the K1 term) against the reference's Z_pred, with torch stand-ins for the C-ABI ops."""
    from conftest import load_golden
    from helpers import build_modules, rel
    from lvae_b200 import utils as U
    g = load_golden(name)
    L = g["mu"].shape[1]
    with emulated_ops():
        cm0, cm1, lik = build_modules(g["lists"], L, g["lengthscale"], g["outputscale"], g["noise"], "cpu")
        t = lambda k: torch.from_numpy(g[k].copy())
        Zv = U.batch_predict_varying_T(L, cm0, cm1, lik, t("x"), t("test_x"), t("mu"), t("z"), 2, float(g["eps"]))
        assert Zv.shape == g["Z_pred"].shape and rel(Zv, g["Z_pred"]) < 1e-6
        if not bool(g["ragged"]):
            P = len(g["offsets"]) - 1
            Z = U.batch_predict(L, cm0, cm1, lik, t("x"), t("test_x"), t("mu"), t("z"), P, int(g["T"]), 2, float(g["eps"]))
            assert rel(Z, g["Z_pred"]) < 1e-6
        perm = torch.randperm(g["x"].shape[0], generator=torch.Generator().manual_seed(1))
        Zp = U.batch_predict_varying_T(L, cm0, cm1, lik, t("x")[perm], t("test_x"), t("mu")[perm], t("z"), 2, float(g["eps"]))
        assert rel(Zp, g["Z_pred"]) < 1e-6
