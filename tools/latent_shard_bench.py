"""Step time of the public API with latent-dimension sharding at the reference's default minibatch (20 subjects), cfg3
(L = 64, M = 256: the per-latent M x M work dominates).  torchrun --nproc-per-node N tools/latent_shard_bench.py"""
import os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
rank, world, lr_ = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import numpy as np
import lvae_b200.elbo_functions as EF
from lvae_b200 import distributed as D, synth
from lvae_b200.training import natural_gradient_step
from helpers import build_modules
cfg, P = os.environ.get("CFG", "cfg3"), 20
b = synth.make_batch(cfg, P=P)
L, M = b.L, b.M
cm0, cm1, lik = build_modules(b.lists, L, np.full((4, L), 2.5), np.full((5, L), 0.69), np.ones(L), dev)
c = lambda t: t.to(dev)
x, z, mu0, lv0 = c(b.x), c(b.z), c(b.mu), c(b.log_v)
st = {"m": c(b.m), "H": c(b.H)}
if world > 1:
    D.enable(shard="latents")
EF.set_error_check("deferred")

def step():
    mu, lv = mu0.clone().requires_grad_(True), lv0.clone().requires_grad_(True)
    kld, gm, gH = EF.minibatch_KLD_upper_bound(cm0, cm1, lik, L, st["m"], st["H"], x, mu, lv, z, 1000, P, b.T, True, 1e-6)
    kld.sum().backward()
    st["m"], st["H"] = natural_gradient_step(st["m"], st["H"], gm, gH, 1e-3)
    cm0.zero_grad(set_to_none=True); cm1.zero_grad(set_to_none=True)

for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(30):
    step()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 30 * 1e3
if rank == 0:
    print(f"{cfg} spb=20 L={L} M={M} ranks={world} shard=latents: {ms:.3f} ms/step", flush=True)
if world > 1:
    D.disable(); dist.destroy_process_group()
