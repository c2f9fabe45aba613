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


class PeerStats:
    """The exchange step over NVLink peer memory instead of an NCCL all-reduce: every rank's reduce kernel writes its
    statistics row into a symmetric-memory buffer (torch.distributed._symmetric_memory, mapped into all ranks of the node),
    one device-side barrier makes the rows visible, and `lvae_peer_sum_f64` sums the peers' rows in rank order — one
    P2P read of (world x 0.9 MB at cfg2) per GPU, bit-identical results on all ranks.  Two buffers alternate between
    steps, so a rank that runs ahead never overwrites a row a slower peer is still reading."""

    def __init__(self, group, numel, device):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        self.numel = int(numel) + (int(numel) & 1)                  # keep both halves 16-byte aligned
        self.buf = symm_mem.empty(2 * self.numel, dtype=torch.float64, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.world = self.hdl.world_size
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.step = 0
        self._arr = [(C.c_uint64 * self.world)(*[p + h * self.numel * 8 for p in self.ptrs]) for h in range(2)]

    def region(self):
        """This step's local destination (the reduce kernel writes here)."""
        h = self.step & 1
        return self.buf[h * self.numel:(h + 1) * self.numel]

    def all_reduce_into(self, out):
        from . import _lib
        h = self.step & 1
        self.hdl.barrier(channel=h)                                 # all ranks' rows of this step are complete and visible
        with torch.cuda.device(out.device):
            _lib.check(_lib.load().lvae_peer_sum_f64(self._arr[h], self.world, out.numel(), _lib.ptr(out),
                                                     _lib.stream_ptr(out.device)), "lvae_peer_sum_f64")
        self.step += 1
        return out


_PEER = {}


def peer_stats(group, numel, device):
    key = (id(group), int(numel), str(device))
    if key not in _PEER:
        _PEER[key] = PeerStats(group, numel, device)
    return _PEER[key]


def enable(group=None, exchange="nccl", shard="subjects", tail="replicated"):
    """Route minibatch_KLD_upper_bound[_iter] through `group` (default: WORLD).
    shard="subjects" (default): every rank passes ITS rows; P_batch / P_in_current_batch remain the GLOBAL minibatch subject
        counts; the statistics row is summed over ranks.  exchange: "nccl" (one all_reduce) | "p2p" (symmetric memory +
        lvae_peer_sum_f64, single node, see PeerStats).
    shard="latents": every rank passes the SAME minibatch and computes latents latent_slice(L, rank, world) of the bound (for
        minibatches too small to split by subject): kld_total comes back all-reduced, grad_m / grad_H all-gathered;
        gradients w.r.t. mu, log_v and the hyper-parameters cover this rank's latent columns — sum them over ranks
        (all_reduce) for the full gradient.
    tail="latents" (with shard="subjects", natural_gradient=True, L divisible by the number of ranks): the subject
        pass stays sharded by subject, but head, tail and natural-gradient update run for L / world latents per rank: the
        statistics rows are reduce-scattered by latent, W = c (G - Kzz^-1) and a = Kzz^-1 m all-gathered before the subject
        pass, kld per latent and the hyper-parameter gradients all-gathered after the tail.  The returned grad_m / grad_H hold
        this rank's latents (zeros elsewhere); training.natural_gradient_step updates those latents and all-gathers the new
        (m, H).  Worth it when the per-latent O(M^3) work is visible next to the rank's share of the subject pass (M >= 128,
        or strong scaling of a fixed minibatch)."""
    if not dist.is_initialized():
        raise RuntimeError("lvae_b200.distributed.enable: torch.distributed is not initialised")
    if exchange not in ("nccl", "p2p"):
        raise ValueError(exchange)
    if shard not in ("subjects", "latents"):
        raise ValueError(shard)
    if tail not in ("replicated", "latents"):
        raise ValueError(tail)
    elbo_functions.set_process_group(group if group is not None else dist.group.WORLD, exchange, shard, tail)


def disable():
    elbo_functions.set_process_group(None)
    _PEER.clear()


def reduce_grads(encoder_params=(), kernel_params=(), group=None):
    """Make the .grad of every parameter the gradient of the GLOBAL loss after backward() through the sharded bound.

    The two shard modes need OPPOSITE reductions, and wrapping the model in DistributedDataParallel (which AVERAGES every
    gradient) is wrong for both:
      shard="subjects": the tail runs on the summed statistics on every rank, so the gradients w.r.t. the kernel
          hyper-parameters, the noise, m and H are already COMPLETE and identical on all ranks — they must NOT be reduced.
          d_mu / d_log_v cover the rank's own rows only, so everything upstream of them (the encoder, and the decoder through
          the rank's share of the reconstruction loss) holds a PARTIAL gradient: SUM over ranks.
      shard="latents": every rank sees all rows but only its latent columns: encoder gradients AND kernel / noise gradients
          are partial: SUM both.
    encoder_params: parameters of the networks (partial in both modes); kernel_params: kernel modules' + likelihood's
    parameters (and m, H when natural_gradient=False).  One flattened all_reduce per group of parameters."""
    group = group if group is not None else elbo_functions._GROUP
    if group is None:
        return
    todo = list(encoder_params)
    if elbo_functions._SHARD == "latents":
        todo += list(kernel_params)
    grads = [p.grad for p in todo if p is not None and p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def all_reduce_stats(stats, group=None):
    """Sum the SVGP sufficient statistics over ranks in place (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
