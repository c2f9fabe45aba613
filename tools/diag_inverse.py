"""Diagnostic: accuracy of Kzz^-1 (this library vs LAPACK vs cuSOLVER) against an extended-precision inverse, on the bench's Kzz."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench, lvae_oracle as orc, lvae_oracle_xp as oxp
from lvae_b200 import ops
dev = torch.device("cuda", 0)
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
for cfg, M in (("cfg2", 60), ("cfg5", 128), ("cfg3", 256)):
    L = 4
    b = bench.make_problem(cfg, 200, 0, 1, L, M)
    k0, k1, nz, _ = bench.oracle_components(b, "cpu", False)
    Kzz = orc.dense(k0, b.z, b.z, L) + 1e-6 * torch.eye(M, dtype=torch.float64)
    # extended precision
    Kx = np.stack([oxp._dense(k0, b.z[l].numpy().astype(oxp.XP), b.z[l].numpy().astype(oxp.XP), l) for l in range(L)]) + oxp.XP(1e-6) * np.eye(M, dtype=oxp.XP)
    Lx, Kix = oxp._spd_inv(Kx)
    Kix64 = Kix.astype(np.float64)
    # FP64 inverse of the FP64-rounded Kzz in extended precision (isolates the inversion from the kernel-entry rounding)
    _, Kix_of64 = oxp._spd_inv(Kzz.numpy().astype(oxp.XP))
    Kix_of64 = Kix_of64.astype(np.float64)
    Lc = torch.linalg.cholesky(Kzz)
    Ki_cpu = torch.cholesky_solve(torch.eye(M, dtype=torch.float64), Lc).numpy()
    Kd = Kzz.to(dev)
    Ki_gpu = torch.cholesky_solve(torch.eye(M, dtype=torch.float64, device=dev), torch.linalg.cholesky(Kd)).cpu().numpy()
    Ki_ours = ops.potri_batched(ops.potrf_batched(Kd)).cpu().numpy()
    print(cfg, M, "cond %.2e" % float(torch.linalg.cond(Kzz).max()))
    for name, K in (("lapack", Ki_cpu), ("cusolver", Ki_gpu), ("ours", Ki_ours)):
        print("   %-9s vs exact-inverse-of-exact-K %.2e   vs exact-inverse-of-fp64-K %.2e   residual |K Ki - I| %.2e" % (
            name, rel(K, Kix64), rel(K, Kix_of64), float(np.abs(Kzz.numpy() @ K - np.eye(M)).max())))
    print("   kernel-entry rounding alone: exact-inverse-of-fp64-K vs exact-inverse-of-exact-K %.2e" % rel(Kix_of64, Kix64))
    m = b.m.numpy()[:, :, 0]
    for name, K in (("lapack", Ki_cpu), ("cusolver", Ki_gpu), ("ours", Ki_ours)):
        a = np.einsum('lij,lj->li', K, m); ax = np.einsum('lij,lj->li', Kix_of64, m)
        print("   %-9s a = Ki m: %.2e" % (name, rel(a, ax)))
