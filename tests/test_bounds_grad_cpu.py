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
