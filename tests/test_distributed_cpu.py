"""CPU (gloo, world_size 2): host-side logic of the subject-sharded path — shard partition and the additivity of the SVGP
sufficient statistics under the all-reduce the GPU path performs between its subject pass and its tail."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden, oracle_components

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_subjects_partitions_exactly():
    from lvae_b200.distributed import shard_rows, shard_subjects
    rng = np.random.default_rng(0)
    for P, world in [(1, 2), (7, 2), (20, 4), (1000, 8), (33, 8)]:
        lens = rng.integers(5, 41, size=P)
        offsets = np.concatenate([[0], np.cumsum(lens)])
        cuts = [shard_subjects(offsets, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == P
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))          # contiguous, disjoint, complete
        rows = [offsets[hi] - offsets[lo] for lo, hi in cuts]
        if P >= 4 * world:
            assert max(rows) - min(rows) <= 2 * 40                          # balanced by rows up to one subject
        lo, hi, loc = shard_rows(offsets, world - 1, world)
        assert hi == offsets[-1] and loc[0] == 0 and loc[-1] == hi - lo


def _partial_stats(g, rows, P_tot, P_b):
    """Per-shard statistics in torch (oracle kernels): S, ng1, and the scalar sum A+Bt+C+D1-F, as the CUDA subject pass
    accumulates them (elbo_functions.py:183-196)."""
    import lvae_oracle as orc
    k0, k1 = oracle_components(g)
    t = lambda k: torch.from_numpy(g[k])
    L, T = g["mu"].shape[1], int(g["T"])
    x, mu, lv, z = t("x")[rows], t("mu")[rows], t("log_v")[rows], t("z")
    P_loc = x.shape[0] // T
    M = z.shape[1]
    with torch.no_grad():
        Kxz = orc.dense(k0, x, z, L)
        Kzz = orc.dense(k0, z, z, L) + float(g["eps"]) * torch.eye(M, dtype=torch.float64)
        Ki = torch.linalg.inv(Kzz)
        xs = x.reshape(P_loc, T, -1).unsqueeze(1).expand(P_loc, L, T, x.shape[1])
        K0s = orc.dense(k0, xs, xs, L).transpose(0, 1)
        Bs = (orc.dense(k1, xs, xs, L) + torch.eye(T, dtype=torch.float64) * t("noise").view(L, 1, 1)).transpose(0, 1)
        Bi = torch.linalg.inv(Bs)
        Kp = Kxz.reshape(L, P_loc, T, M)
        S = (Kp.transpose(-1, -2) @ Bi @ Kp).sum(1)
        r = ((Kxz @ Ki) @ t("m")).squeeze(-1) - mu.T
        rs = r.reshape(L, P_loc, T, 1)
        A = (rs.transpose(2, 3) @ Bi @ rs).sum()
        Bt = (torch.diagonal(Bi, dim1=-1, dim2=-2).reshape(L, -1) * torch.exp(lv.T)).sum()
        C = torch.logdet(Bs).sum()
        D1 = (Bi * K0s).sum()
        F = lv.sum()
        ng1 = (Kp.transpose(-1, -2) @ (Bi @ mu.T.reshape(L, P_loc, T, 1))).sum(1)
    return torch.cat([S.reshape(-1), ng1.reshape(-1), (A + Bt + C + D1 - F).reshape(1)])


def _worker(rank, world, port, case, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lvae_b200.distributed import all_reduce_stats, shard_rows
    g = load_golden(case)
    lo, hi, _ = shard_rows(g["offsets"], rank, world)
    stats = _partial_stats(g, slice(lo, hi), int(g["P_tot"]), len(g["offsets"]) - 1)
    all_reduce_stats(stats)
    if rank == 0:
        torch.save(stats, out)
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_statistics_sum_to_full_batch(tmp_path):
    import lvae_oracle as orc
    case = "cfg2_small"
    out = str(tmp_path / "stats.pt")
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 500, case, out), nprocs=2, join=True)
    reduced = torch.load(out)
    g = load_golden(case)
    full = _partial_stats(g, slice(0, g["x"].shape[0]), int(g["P_tot"]), len(g["offsets"]) - 1)
    assert torch.allclose(reduced, full, rtol=1e-12, atol=1e-9)
    # and the tail on the reduced statistics reproduces the reference's bound
    t = lambda k: torch.from_numpy(g[k])
    k0, k1 = oracle_components(g)
    L, M, T = g["mu"].shape[1], g["z"].shape[1], int(g["T"])
    P_b, P_tot = len(g["offsets"]) - 1, int(g["P_tot"])
    with torch.no_grad():
        S = reduced[:L * M * M].reshape(L, M, M)
        scal = reduced[-1]
        Kzz = orc.dense(k0, t("z"), t("z"), L) + 1e-6 * torch.eye(M, dtype=torch.float64)
        Ki = torch.linalg.inv(Kzz)
        H, m = t("H"), t("m")
        G = Ki @ H @ Ki
        D2, E = (S * Ki).sum(), (G.transpose(-1, -2) * S).sum()
        kl = 0.5 * ((Ki * H.transpose(-1, -2)).sum() + (m * (Ki @ m)).sum() - L * M + torch.logdet(Kzz).sum() - torch.logdet(H).sum())
        kld = P_tot / P_b * 0.5 * (scal - D2 + E) + kl - L * P_tot * T / 2
    assert abs(kld.item() - float(g["kld"])) <= 1e-7 * abs(float(g["kld"]))


def test_latent_slices_partition_the_latent_dimensions():
    """shard="latents": contiguous, balanced blocks that cover 0..L-1 exactly once (host logic of distributed.enable)."""
    from lvae_b200.elbo_functions import latent_slice
    for L, world in [(32, 8), (5, 2), (7, 3), (64, 8), (3, 3), (10, 4)]:
        blocks = [latent_slice(L, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == L
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1


def _worker_reduce_grads(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lvae_b200.elbo_functions as EF
    from lvae_b200.distributed import reduce_grads
    res = {}
    for shard in ("subjects", "latents"):
        EF.set_process_group(dist.group.WORLD, "nccl", shard)
        enc = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2, 2))]
        ker = [torch.nn.Parameter(torch.zeros(4))]
        enc[0].grad = torch.full((3,), float(rank + 1))
        enc[1].grad = torch.full((2, 2), 10.0 * (rank + 1))
        ker[0].grad = torch.full((4,), 100.0 * (rank + 1))
        reduce_grads(enc, ker)
        res[shard] = (enc[0].grad.clone(), enc[1].grad.clone(), ker[0].grad.clone())
    EF.set_process_group(None)
    if rank == 1:
        torch.save(res, out)
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_reduce_grads_applies_the_reduction_each_shard_mode_needs(tmp_path):
    """subjects: encoder gradients are partial (summed), kernel gradients complete (untouched); latents: both summed."""
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker_reduce_grads, args=(2, 29100 + os.getpid() % 300, out), nprocs=2, join=True)
    res = torch.load(out)
    e0, e1, k = res["subjects"]
    assert torch.equal(e0, torch.full((3,), 3.0)) and torch.equal(e1, torch.full((2, 2), 30.0))
    assert torch.equal(k, torch.full((4,), 200.0))                 # rank 1's own, complete gradient: not reduced
    e0, e1, k = res["latents"]
    assert torch.equal(e0, torch.full((3,), 3.0)) and torch.equal(k, torch.full((4,), 300.0))


def test_hensman_training_refuses_to_run_sharded():
    import lvae_b200.elbo_functions as EF
    from lvae_b200.training import hensman_training
    EF.set_process_group(object())
    try:
        with pytest.raises(RuntimeError, match="single-process"):
            hensman_training(None, None, 1, [], None, 'GPapprox_closed', 1, 2, None, None, None, None, None, None, 1, 1, False,
                             6, 1.0, 2, 'mse')
    finally:
        EF.set_process_group(None)
