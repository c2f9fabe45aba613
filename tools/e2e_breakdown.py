"""Where does the end-to-end (host buffers -> public API -> host results) step spend its time?  Run on the GPU box."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lvae_b200.elbo_functions as EF
from lvae_b200 import synth
from lvae_b200.training import natural_gradient_step
from helpers import build_modules
import numpy as np

import os
P, L, M = int(os.environ.get("E2E_P", "1000")), 32, 60
b = synth.make_batch("cfg2", P=P, L=L, M=M)
dev = "cuda"
cm0, cm1, lik = build_modules(b.lists, L, np.full((4, L), 2.5), np.full((5, L), 0.69), np.ones(L), dev)
hx, hmu, hlv = b.x.pin_memory(), b.mu.pin_memory(), b.log_v.pin_memory()
z, m, H = b.z.cuda(), b.m.cuda(), b.H.cuda()
out_mu = torch.empty_like(b.mu).pin_memory(); out_lv = torch.empty_like(b.log_v).pin_memory(); out_k = torch.empty(1, dtype=torch.float64).pin_memory()

def step(sync_marks=None):
    t = [time.perf_counter()]
    def mark():
        if sync_marks is not None:
            torch.cuda.synchronize(); t.append(time.perf_counter())
    xd = hx.to(dev, non_blocking=True); mud = hmu.to(dev, non_blocking=True).requires_grad_(True); lvd = hlv.to(dev, non_blocking=True).requires_grad_(True)
    mark()
    kld, gm, gH = EF.minibatch_KLD_upper_bound(cm0, cm1, lik, L, m, H, xd, mud, lvd, z, P, P, 20, True, 1e-6)
    mark()
    kld.backward()
    mark()
    m2, H2 = natural_gradient_step(m, H, gm, gH, 1e-3)
    mark()
    out_mu.copy_(mud.grad, non_blocking=True); out_lv.copy_(lvd.grad, non_blocking=True); out_k.copy_(kld.detach().reshape(1), non_blocking=True)
    mark()
    if sync_marks is not None:
        sync_marks.append([b_ - a_ for a_, b_ in zip(t, t[1:])])

for _ in range(5): step()
torch.cuda.synchronize()
marks = []
for _ in range(10): step(marks)
print("synced phases ms [h2d, fwd, bwd, ng, d2h]:", (np.array(marks).mean(0) * 1e3).round(3))
for mode in ("immediate", "deferred", "immediate", "deferred"):
    EF.set_error_check(mode)
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): step()
    e1.record()
    torch.cuda.synchronize(); print(mode, "async e2e ms/step: wall", (time.perf_counter() - t0) / 20 * 1e3, "events", e0.elapsed_time(e1) / 20)
    EF.check_errors()
EF.set_error_check("immediate")
# host-only cost of the forward call (no GPU wait): time the python until return
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(35)
