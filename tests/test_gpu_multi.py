"""GPU, >= 2 devices (gpurun --gpus 2): subjects sharded over two ranks, statistics all-reduced over NCCL — bound, natural
gradients and hyper-gradients must equal the single-GPU result; d_mu rows follow their owner."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, case, out_dir, exchange):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import lvae_b200.elbo_functions as EF
    from lvae_b200 import distributed as D
    from helpers import build_modules, constrained_param_grads
    g = load_golden(case)
    dev = f"cuda:{rank}"
    L = g["mu"].shape[1]
    lo, hi, _ = D.shard_rows(g["offsets"], rank, world)
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    cm0, cm1, lik = build_modules(g["lists"], L, g["lengthscale"], g["outputscale"], g["noise"], dev)
    mu = t("mu")[lo:hi].clone().requires_grad_(True)
    lv = t("log_v")[lo:hi].clone().requires_grad_(True)
    P_b = len(g["offsets"]) - 1
    D.enable(exchange=exchange)
    if bool(g["ragged"]):
        kld, gm, gH = EF.minibatch_KLD_upper_bound_iter(cm0, cm1, lik, L, t("m"), t("H"), t("x")[lo:hi], mu, lv, t("z"),
                                                        int(g["P_tot"]), P_b, int(g["N_tot"]), True, 2, float(g["eps"]))
    else:
        kld, gm, gH = EF.minibatch_KLD_upper_bound(cm0, cm1, lik, L, t("m"), t("H"), t("x")[lo:hi], mu, lv, t("z"),
                                                   int(g["P_tot"]), P_b, int(g["T"]), True, float(g["eps"]))
    kld.sum().backward()
    D.disable()
    torch.save(dict(kld=float(kld.sum().item()), gm=gm.cpu(), gH=gH.cpu(), d_mu=mu.grad.cpu(), lo=lo, hi=hi,
                    d_hyper=constrained_param_grads(cm0, cm1, lik)), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
@pytest.mark.parametrize("case", ["cfg2_small", "cfg4_ragged", "cfg3_small"])
def test_two_rank_sharding_matches_reference(case, exchange, tmp_path):
    """exchange = "p2p": the statistics row is summed by lvae_peer_sum_f64 over symmetric memory instead of NCCL."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from helpers import golden_hyper_vector, rel
    mp.spawn(_worker, args=(2, 29700 + os.getpid() % 200, case, str(tmp_path), exchange), nprocs=2, join=True)
    g = load_golden(case)
    outs = [torch.load(os.path.join(str(tmp_path), f"r{r}.pt"), weights_only=False) for r in range(2)]
    for o in outs:
        assert abs(o["kld"] - float(g["kld"])) <= 1e-6 * abs(float(g["kld"]))
        assert rel(o["gm"], g["grad_m"]) < 1e-6 and rel(o["gH"], g["grad_H"]) < 1e-6
        ref = golden_hyper_vector(g)
        assert np.abs(o["d_hyper"] - ref).max() <= 1e-6 * np.abs(ref).max()
        assert rel(o["d_mu"], g["d_mu"][o["lo"]:o["hi"]]) < 1e-6
    assert outs[0]["hi"] == outs[1]["lo"]
    if exchange == "p2p":        # rank-ordered sums: the replicated tail gives bit-identical results on both ranks
        assert outs[0]["kld"] == outs[1]["kld"] and torch.equal(outs[0]["gH"], outs[1]["gH"])


def _worker_latent(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import lvae_b200.elbo_functions as EF
    from lvae_b200 import distributed as D
    from helpers import build_modules
    g = load_golden(case)
    dev = f"cuda:{rank}"
    L = g["mu"].shape[1]
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    cm0, cm1, lik = build_modules(g["lists"], L, g["lengthscale"], g["outputscale"], g["noise"], dev)
    mu = t("mu").clone().requires_grad_(True)
    lv = t("log_v").clone().requires_grad_(True)
    P_b = len(g["offsets"]) - 1
    D.enable(shard="latents")
    if bool(g["ragged"]):
        kld, gm, gH = EF.minibatch_KLD_upper_bound_iter(cm0, cm1, lik, L, t("m"), t("H"), t("x"), mu, lv, t("z"),
                                                        int(g["P_tot"]), P_b, int(g["N_tot"]), True, 2, float(g["eps"]))
    else:
        kld, gm, gH = EF.minibatch_KLD_upper_bound(cm0, cm1, lik, L, t("m"), t("H"), t("x"), mu, lv, t("z"),
                                                   int(g["P_tot"]), P_b, int(g["T"]), True, float(g["eps"]))
    kld.sum().backward()
    D.disable()
    l0, l1 = EF.latent_slice(L, rank, world)
    own = torch.zeros(L, dtype=torch.bool)
    own[l0:l1] = True
    local_only = bool((mu.grad[:, ~own.to(dev)] == 0).all())        # gradients only reach this rank's latent columns
    dist.all_reduce(mu.grad)
    dist.all_reduce(lv.grad)
    for p in list(cm0.parameters()) + list(cm1.parameters()) + list(lik.parameters()):
        dist.all_reduce(p.grad)
    from helpers import constrained_param_grads
    torch.save(dict(kld=float(kld.sum().item()), gm=gm.cpu(), gH=gH.cpu(), d_mu=mu.grad.cpu(), d_lv=lv.grad.cpu(),
                    local_only=local_only, d_hyper=constrained_param_grads(cm0, cm1, lik)), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["cfg2_small", "cfg4_ragged", "cfg3_small"])
def test_two_rank_latent_sharding_matches_reference(case, tmp_path):
    """shard="latents": both ranks see the whole minibatch, each computes its latent dimensions; kld all-reduced, natural
    gradients all-gathered, parameter gradients summed over ranks equal the single-GPU / reference values."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from helpers import golden_hyper_vector, rel
    mp.spawn(_worker_latent, args=(2, 29500 + os.getpid() % 200, case, str(tmp_path)), nprocs=2, join=True)
    g = load_golden(case)
    ref = golden_hyper_vector(g)
    for r in range(2):
        o = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"), weights_only=False)
        assert o["local_only"]
        assert abs(o["kld"] - float(g["kld"])) <= 1e-6 * abs(float(g["kld"]))
        assert rel(o["gm"], g["grad_m"]) < 1e-6 and rel(o["gH"], g["grad_H"]) < 1e-6
        assert rel(o["d_mu"], g["d_mu"]) < 1e-6 and rel(o["d_lv"], g["d_log_v"]) < 1e-6
        assert np.abs(o["d_hyper"] - ref).max() <= 1e-6 * np.abs(ref).max()


def _worker_latent_tail(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import lvae_b200.elbo_functions as EF
    from lvae_b200 import distributed as D
    from lvae_b200.training import natural_gradient_step
    from helpers import build_modules, constrained_param_grads
    g = load_golden(case)
    dev = f"cuda:{rank}"
    L = g["mu"].shape[1]
    lo, hi, _ = D.shard_rows(g["offsets"], rank, world)
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    cm0, cm1, lik = build_modules(g["lists"], L, g["lengthscale"], g["outputscale"], g["noise"], dev)
    mu = t("mu")[lo:hi].clone().requires_grad_(True)
    lv = t("log_v")[lo:hi].clone().requires_grad_(True)
    P_b = len(g["offsets"]) - 1
    m, H = t("m"), t("H")
    D.enable(tail="latents")
    kld, gm, gH = EF.minibatch_KLD_upper_bound(cm0, cm1, lik, L, m, H, t("x")[lo:hi], mu, lv, t("z"),
                                               int(g["P_tot"]), P_b, int(g["T"]), True, float(g["eps"]))
    kld.sum().backward()
    m1, H1 = natural_gradient_step(m, H, gm, gH, 0.05)
    l0, l1 = L * rank // world, L * (rank + 1) // world
    foreign_zero = bool((gH[:l0] == 0).all() and (gH[l1:] == 0).all())      # only this rank's latents are filled
    gmf, gHf = gm.clone(), gH.clone()
    dist.all_reduce(gmf)                                                        # zeros elsewhere: the sum is the full gradient
    dist.all_reduce(gHf)
    D.disable()
    torch.save(dict(kld=float(kld.sum().item()), gm=gmf.cpu(), gH=gHf.cpu(), d_mu=mu.grad.cpu(), lo=lo, hi=hi, m1=m1.cpu(),
                    H1=H1.cpu(), foreign_zero=foreign_zero, d_hyper=constrained_param_grads(cm0, cm1, lik)),
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["cfg2_small", "cfg3_small"])
def test_two_rank_latent_sharded_tail_matches_reference(case, tmp_path):
    """distributed.enable(tail="latents"): subject pass sharded by subject, head / tail / natural-gradient update by latent
    (reduce-scatter of the statistics by latent, all-gather of W, a and of the new (m, H)); results equal the reference's and
    the updated (m, H) equal the oracle's natural-gradient step on both ranks."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import lvae_oracle as orc
    from helpers import golden_hyper_vector, rel
    mp.spawn(_worker_latent_tail, args=(2, 29300 + os.getpid() % 200, case, str(tmp_path)), nprocs=2, join=True)
    g = load_golden(case)
    ref = golden_hyper_vector(g)
    t = lambda k: torch.from_numpy(g[k])
    m_ref, H_ref = orc.ng_step(t("m"), t("H"), t("grad_m"), t("grad_H"), 0.05)
    for r in range(2):
        o = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"), weights_only=False)
        assert o["foreign_zero"]
        assert abs(o["kld"] - float(g["kld"])) <= 1e-6 * abs(float(g["kld"]))
        assert rel(o["gm"], g["grad_m"]) < 1e-6 and rel(o["gH"], g["grad_H"]) < 1e-6
        assert np.abs(o["d_hyper"] - ref).max() <= 1e-6 * np.abs(ref).max()
        assert rel(o["d_mu"], g["d_mu"][o["lo"]:o["hi"]]) < 1e-6
        assert rel(o["m1"], m_ref) < 1e-6 and rel(o["H1"], H_ref) < 1e-6
