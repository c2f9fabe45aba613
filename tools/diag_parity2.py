"""Diagnostic: CUDA path vs oracle on the GPU vs oracle on the CPU, growing P; where do d_log_v / d_mu rows differ?"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lvae_b200 import ops
from lvae_b200.spec import build_structure, flatten

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
os.environ["LVAE_BENCH_HYPERS"] = os.environ.get("LVAE_BENCH_HYPERS", "default")
path = int(os.environ.get("DIAG_PATH", "0"))
LL = int(os.environ.get("DIAG_L", "2"))
for P in [int(v) for v in os.environ.get("DIAG_P", "40,100,200,400,1000").split(",")]:
    L, M = LL, 60
    b = bench.make_problem("cfg2", P, 0, 1, L, M)
    cm0, cm1, lik = bench.build_modules(b, dev)
    st, ls, os_ = build_structure(flatten(cm0), flatten(cm1), L, device=dev)
    noise = lik.noise.detach().reshape(L).contiguous()
    Tl = np.diff(b.offsets)
    call = ops.KldCall(st, L, M, 6, P, b.N, int(Tl.max()), int((Tl * Tl).sum()), dev, natural_gradient=True, path=path)
    d = lambda t: t.to(dev)
    offs = torch.from_numpy(b.offsets).to(torch.int32).to(dev)
    call.bind(d(b.x), offs, d(b.mu), d(b.log_v), d(b.z), d(b.m).view(L, M), d(b.H), ls.detach(), os_.detach(), noise, 1.0,
              L * P * 20 / 2, 1e-6)
    call.run()
    torch.cuda.synchronize()
    call.raise_on_info()
    with torch.device(dev):
        rg = bench.oracle_step_fn(b, P, device=dev)(update=False)
    rc = bench.oracle_step_fn(b, P, device="cpu")(update=False)
    ours = dict(kld=call.kld_per_latent.sum(), grad_m=call.grad_m, grad_H=call.grad_H, d_mu=call.d_mu, d_log_v=call.d_log_v,
                d_hyper=call.d_hyper)
    for k in ours:
        print("L", L, "path", path, "PREP", os.environ.get("LVAE_PREP"), P, k, "ours-vs-cpu %.2e  gpu-vs-cpu %.2e  ours-vs-gpu %.2e" % (bench.rel_err(ours[k].cpu(), rc[k]),
              bench.rel_err(rg[k].cpu(), rc[k]), bench.rel_err(ours[k], rg[k])), flush=True)
    for name, a in (("ours", ours["d_log_v"].cpu()), ("gpu", rg["d_log_v"].cpu())):
        bad = ((a - rc["d_log_v"]).abs().max(dim=1).values > 1e-6).nonzero().reshape(-1)
        print(P, name, "bad d_log_v rows:", bad.numel(), "first", bad[:8].tolist(), "subjects", sorted(set((bad // 20).tolist()))[:10], flush=True)
