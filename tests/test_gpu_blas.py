"""GPU parity of the batched FP64 building blocks (DMMA GEMM, blocked Cholesky / SPD inverse for 64 < n <= 256) against
torch FP64 (cuBLAS / cuSOLVER).  Tolerance 1e-12 relative for products, 1e-9 for inverses of moderately conditioned SPD
matrices (these kernels replace torch.matmul / torch.cholesky / cholesky_solve of elbo_functions.py:176-186,194)."""
import pytest
import torch

from lvae_b200 import ops

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
# the last three have few output tiles and a long k: the ABI entry splits the k range over CTAs and sums the parts
@pytest.mark.parametrize("m,n,k", [(128, 64, 16), (200, 72, 61), (7, 5, 3), (256, 256, 256), (130, 190, 1000),
                                   (60, 1, 20000), (60, 60, 20001), (7, 5, 5000)])
def test_gemm_matches_torch(ta, tb, m, n, k):
    torch.manual_seed(m * 7 + n * 3 + k)
    batch = 3
    A = torch.randn(batch, *((k, m) if ta else (m, k)), dtype=torch.float64, device="cuda")
    B = torch.randn(batch, *((n, k) if tb else (k, n)), dtype=torch.float64, device="cuda")
    C0 = torch.randn(batch, m, n, dtype=torch.float64, device="cuda")
    ref = 0.7 * (A.transpose(1, 2) if ta else A) @ (B.transpose(1, 2) if tb else B) - 1.3 * C0
    out = ops.gemm_batched(A, B, ta, tb, alpha=0.7, beta=-1.3, C=C0.clone())
    assert rel(out, ref) < 1e-12


@pytest.mark.parametrize("n,k", [(72, 500), (128, 33), (256, 2000), (60, 20000), (256, 4100)])
def test_gemm_syrk_lower_mirror(n, k):
    torch.manual_seed(n + k)
    U = torch.randn(2, k, n, dtype=torch.float64, device="cuda")
    out = ops.gemm_batched(U, U, True, False, flags=3)
    ref = U.transpose(1, 2) @ U
    assert rel(out, ref) < 1e-12
    assert torch.equal(out, out.transpose(1, 2))
    low = ops.gemm_batched(U, U, True, False, flags=1)
    assert torch.equal(torch.tril(low), torch.tril(out)) and float(torch.triu(low, 1).abs().max()) == 0.0


@pytest.mark.parametrize("n", [65, 72, 128, 200, 256])
def test_potrf_potri_big(n):
    torch.manual_seed(n)
    batch = 5
    A = torch.randn(batch, n, 2 * n, dtype=torch.float64, device="cuda")
    A = A @ A.transpose(1, 2) / (2 * n) + 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")
    Lc = ops.potrf_batched(A)
    ref = torch.linalg.cholesky(A)
    assert rel(Lc, ref) < 1e-11
    assert float(torch.triu(Lc, 1).abs().max()) == 0.0
    inv = ops.potri_batched(Lc)
    assert rel(inv, torch.linalg.inv(A)) < 1e-9
    assert rel(inv @ A, torch.eye(n, dtype=torch.float64, device="cuda").expand(batch, n, n)) < 1e-9


def test_potrf_big_flags_non_pd():
    A = torch.eye(100, dtype=torch.float64, device="cuda").repeat(3, 1, 1)
    A[1, 70, 70] = -1.0
    with pytest.raises(RuntimeError):
        ops.potrf_batched(A)
