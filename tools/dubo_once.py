"""One warm-up and one profiled validation_dubo forward + backward at cfg2's shape (for ncu launch lists)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lvae_b200 import synth  # noqa: E402
from lvae_b200.constraints import GreaterThan  # noqa: E402
from lvae_b200.kernel_gen import generate_kernel_batched  # noqa: E402
from lvae_b200.likelihoods import GaussianLikelihood  # noqa: E402
from lvae_b200.validation import validation_dubo  # noqa: E402

b = synth.make_batch("cfg2", P=int(os.environ.get("P", 1000)))
L, T, P = b.L, b.T, b.P
cm0, cm1 = generate_kernel_batched(L, **b.lists, id_covariate=2)
cm0, cm1 = cm0.double().cuda(), cm1.double().cuda()
lik = GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=GreaterThan(1e-8)).double().cuda()
x, z = b.x.cuda(), b.z.cuda()
mu, lv = b.mu.cuda().requires_grad_(True), b.log_v.cuda().requires_grad_(True)
for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    validation_dubo(L, cm0, cm1, lik, x, mu, lv, z, P, T, 1e-6).sum().backward()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
