"""CPU study (torch stand-ins of tests/ops_emulation.py): how far the DUBO hyper-parameter gradients of mathematically
equivalent FP64 inverse algorithms (cholesky_inverse | linalg.inv | cholesky_solve(I)) drift apart at cfg2 scale, where
Kzz = K0(Z,Z) + 1e-6 I has duplicated rows (cond ~ 1e8).  Result (P=300, L=4, M=60): value 3e-11, d mu / d log v 2e-9, K1 and noise
gradients <= 1e-6, K0 outputscale / lengthscale gradients 1e-5 .. 2e-4 — the floor any FP64 implementation of this bound has."""
import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import ops_emulation as emu
from lvae_b200 import synth
from lvae_b200.constraints import GreaterThan
from lvae_b200.kernel_gen import generate_kernel_batched
from lvae_b200.likelihoods import GaussianLikelihood
from lvae_b200.validation import validation_dubo
P, L = 300, 4
b = synth.make_batch("cfg2", P=P, L=L)
T = b.T
def run(variant):
    if variant == "inv":
        emu.potri_batched = lambda Lc: torch.linalg.inv(Lc.detach() @ Lc.detach().transpose(-1,-2))
    elif variant == "solve":
        def f(Lc):
            I = torch.eye(Lc.shape[-1], dtype=Lc.dtype).expand_as(Lc)
            return torch.cholesky_solve(I, Lc.detach())
        emu.potri_batched = f
    torch.manual_seed(0)
    cm0, cm1 = generate_kernel_batched(L, **b.lists, id_covariate=2)
    cm0, cm1 = cm0.double(), cm1.double()
    lik = GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=GreaterThan(1e-8)).double()
    mu, lv = b.mu.clone().requires_grad_(True), b.log_v.clone().requires_grad_(True)
    params = [mu, lv] + [p for m in (cm0, cm1, lik) for p in m.parameters()]
    with emu.emulated_ops():
        v = validation_dubo(L, cm0, cm1, lik, b.x, mu, lv, b.z, P, T, 1e-6).sum()
        v.backward()
    return float(v), [p.grad.clone() for p in params], [n for m in (cm0,cm1,lik) for n,_ in m.named_parameters()]
v0, g0, names = run("cholinv")
for var in ("inv", "solve"):
    v1, g1, _ = run(var)
    print(var, "value rel", abs(v1-v0)/abs(v0))
    for i,(a,c) in enumerate(zip(g0,g1)):
        nm = (["mu","log_v"]+names)[i]
        print(f"   {nm:45s} rel diff {float((a-c).abs().max()/c.abs().max()):.2e}  |g| {float(c.abs().max()):.3e}")
