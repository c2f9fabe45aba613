"""Per-op device time of the pieces validation_dubo / batch_predict are composed of, at cfg2's shape (CUDA events)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    from lvae_b200 import ops, synth
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200.spec import build_structure, flatten
    P = int(os.environ.get("P", 1000))
    b = synth.make_batch("cfg2", P=P)
    L, T, M = b.L, b.T, b.M
    N = P * T
    cm0, cm1 = generate_kernel_batched(L, **b.lists, id_covariate=2)
    st, ls, os_ = build_structure(flatten(cm0.double().cuda()), flatten(cm1.double().cuda()), L, device="cuda")
    ls, os_ = ls.detach(), os_.detach()
    x, z = b.x.cuda(), b.z.cuda()
    noise = torch.ones(L, dtype=torch.float64, device="cuda")
    off = torch.arange(0, N + 1, T, dtype=torch.int32, device="cuda")
    K0xz = ops.kernel_dense(st, ls, os_, x, z, "k0")
    K0zz = ops.kernel_dense(st, ls, os_, z, z, "k0") + 1e-6 * torch.eye(M, dtype=torch.float64, device="cuda")
    B_st = ops.kernel_blocks(st, ls, os_, x, off, P * T * T, "k1", diag_add=noise).reshape(L * P, T, T)
    LB = ops.potrf_batched(B_st)
    iB = ops.potri_batched(LB)
    iBK = torch.bmm(iB, K0xz.reshape(L * P, T, M)).reshape(L, N, M)
    G = torch.randn_like(K0xz)
    Gb = torch.randn(L, P * T * T, dtype=torch.float64, device="cuda")
    r = {}
    r["kernel_dense K0xz"] = timed(lambda: ops.kernel_dense(st, ls, os_, x, z, "k0"))
    r["kernel_dense K0zz"] = timed(lambda: ops.kernel_dense(st, ls, os_, z, z, "k0"))
    r["kernel_blocks k0"] = timed(lambda: ops.kernel_blocks(st, ls, os_, x, off, P * T * T, "k0"))
    r["kernel_blocks k1"] = timed(lambda: ops.kernel_blocks(st, ls, os_, x, off, P * T * T, "k1", diag_add=noise))
    r["potrf blocks (no check)"] = timed(lambda: ops.potrf_batched(B_st, check_info=False))
    r["potrf blocks (check)"] = timed(lambda: ops.potrf_batched(B_st))
    r["potri blocks"] = timed(lambda: ops.potri_batched(LB))
    r["torch cholesky blocks"] = timed(lambda: torch.linalg.cholesky(B_st))
    r["torch cholesky_inverse blocks"] = timed(lambda: torch.cholesky_inverse(LB))
    r["potrf Kzz"] = timed(lambda: ops.potrf_batched(K0zz, check_info=False))
    r["potri Kzz"] = timed(lambda: ops.potri_batched(K0zz))
    r["bmm iB K0xz"] = timed(lambda: torch.bmm(iB, K0xz.reshape(L * P, T, M)))
    r["gemm S (TN, k=N)"] = timed(lambda: ops.gemm_batched(K0xz, iBK, trans_a=True))
    r["gemm SD (TN, syrk)"] = timed(lambda: ops.gemm_batched(iBK, iBK, trans_a=True, flags=3))
    r["torch bmm S"] = timed(lambda: torch.bmm(K0xz.transpose(1, 2), iBK))
    r["kernel_dense_bwd K0xz"] = timed(lambda: ops.kernel_dense_bwd(st, ls, os_, x, z, G, "k0"))
    r["kernel_blocks_bwd k1"] = timed(lambda: ops.kernel_blocks_bwd(st, ls, os_, x, off, Gb, "k1", want_diag=True))
    for k, v in r.items():
        print(f"{k:34s} {v:9.3f} ms")


if __name__ == "__main__":
    main()
