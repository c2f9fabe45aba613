"""Multi-GPU layout of the GP-prior ELBO path: one process per GPU, the subjects of a minibatch sharded across ranks.

Every batch-dependent term of the bound is a sum over subjects (S, ng1, K_xz^T B^-1 r, A, Bt, C, D1, F and the subject
part of the hyper-parameter adjoints; elbo_functions.py:190-196, 278-288), so the only exchange step is ONE all-reduce of
the per-latent statistics row (8 * L * (M^2 + 2M + 8 + n_hyper) bytes) between the subject pass and the per-latent tail.
Inducing points, (m, H) and hyper-parameters are replicated; d_mu / d_log_v stay local to the rank that owns the rows;
the tail is computed redundantly on every rank, so kld, grad_m, grad_H and the hyper-parameter gradients are identical
everywhere without a second collective.  The reference has no distributed code — this is new capability.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import elbo_functions


def shard_subjects(offsets, rank, world_size):
    """Contiguous slice of whole subjects for `rank`, balanced by ROW count (ragged T): returns (p_lo, p_hi).
    offsets: int array [P+1] of row offsets in the sampler's global subject order; every subject lands on one rank."""
    offsets = np.asarray(offsets, dtype=np.int64)
    P = len(offsets) - 1
    total = int(offsets[-1])
    targets = [(total * r) // world_size for r in range(world_size + 1)]
    cuts = [int(np.searchsorted(offsets, t, side="left")) for t in targets]
    cuts[0], cuts[-1] = 0, P
    for r in range(1, world_size + 1):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts[rank], cuts[rank + 1]


def shard_rows(offsets, rank, world_size):
    """(row_lo, row_hi, local_offsets) of the rank's shard."""
    offsets = np.asarray(offsets, dtype=np.int64)
    p_lo, p_hi = shard_subjects(offsets, rank, world_size)
    return int(offsets[p_lo]), int(offsets[p_hi]), offsets[p_lo:p_hi + 1] - offsets[p_lo]


def enable(group=None):
    """Route the statistics all-reduce of minibatch_KLD_upper_bound[_iter] through `group` (default: WORLD).
    After this, every rank passes ITS rows; P_batch / P_in_current_batch remain the GLOBAL minibatch subject counts."""
    if not dist.is_initialized():
        raise RuntimeError("lvae_b200.distributed.enable: torch.distributed is not initialised")
    elbo_functions.set_process_group(group if group is not None else dist.group.WORLD)


def disable():
    elbo_functions.set_process_group(None)


def all_reduce_stats(stats, group=None):
    """Sum the SVGP sufficient statistics over ranks in place (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
