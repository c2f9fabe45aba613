"""Drop-in for the reference's elbo_functions.py on the Hensman minibatch path (elbo_functions.py:144-307).

`minibatch_KLD_upper_bound` and `minibatch_KLD_upper_bound_iter` keep the reference's call signatures and return
`(kld_total, grad_m, grad_H)`.  The arithmetic runs in liblvae_b200.so: one head kernel (per-latent M x M work), the
per-subject pass (kernel blocks built from covariates, batched Cholesky/inverse of the T x T blocks, S and the A..F partial
sums, fused with the reverse pass), an optional NCCL all-reduce of the sufficient statistics when the subjects are
sharded across GPUs, and one tail kernel.  `kld_total` is an autograd node w.r.t. mu, log_v, every kernel
hyper-parameter, the likelihood noise and (natural_gradient=False) m and H; its gradients are produced by the same
launches (closed-form adjoints) and scaled by the incoming gradient in backward.
"""
import collections
import sys

import numpy as np
import torch

from . import ops
from ._lib import MAX_M, MAX_T
from .spec import build_structure, flatten

_GROUP = None        # torch.distributed process group over which the minibatch subjects are sharded (None = 1 GPU)
_PATH = 0            # 0 auto, 1 generic kernels, 2 fused DMMA kernel
_CHECK = "immediate"  # Cholesky-failure check: "immediate" (one device sync per call, like torch.cholesky) | "deferred"
_PENDING = []        # KldCall objects whose info flags have not been read yet (deferred mode)


_EXCHANGE = "nccl"   # how the statistics row is summed over ranks: "nccl" all-reduce | "p2p" (distributed.PeerStats)
_SHARD = "subjects"  # "subjects": every rank holds its rows of the minibatch | "latents": every rank holds its latent dimensions
_TAIL = "replicated"  # shard="subjects" only: "replicated" (every rank runs head / tail for all latents) | "latents"


def set_process_group(group, exchange="nccl", shard="subjects", tail="replicated"):
    """shard="subjects": every rank passes ITS rows; P_batch / P_in_current_batch stay GLOBAL minibatch subject counts.
        tail="latents": head, tail and natural-gradient update are additionally sharded by latent (ops.LatentTailKldCall).
    shard="latents": every rank passes the SAME full minibatch and computes the bound for its own slice of the latent
    dimensions (no statistics exchange: the latent dimensions are independent); see distributed.enable."""
    global _GROUP, _EXCHANGE, _SHARD, _TAIL
    _GROUP, _EXCHANGE, _SHARD, _TAIL = group, exchange, shard, tail
    _POOL.clear()            # prepared calls may be bound to the previous group (and hold GBs of scratch)


def exchange_stats(call, group, exchange):
    """The one exchange step of the sharded path, between the subject pass and the tail (call.subjects() included)."""
    if group is None or isinstance(call, ops.LatentTailKldCall):      # the latter reduce-scatters by latent itself
        call.subjects()
    elif exchange == "p2p":
        from . import distributed
        ps = distributed.peer_stats(group, call.stats.numel(), call.device)
        call.set_stats(ps.region())                 # the reduce kernel writes straight into symmetric memory
        call.subjects()
        call.set_stats(call.stats)
        ps.all_reduce_into(call.stats)
    else:
        call.subjects()
        torch.distributed.all_reduce(call.stats, group=group)      # SVGP sufficient statistics over NVLink


def set_kernel_path(path):
    global _PATH
    _PATH = int(path)


def set_error_check(mode):
    """"immediate": raise inside the call (costs one device->host sync per step, as torch.cholesky does in the reference).
    "deferred": never block; a failed factorisation makes kld NaN and raises at the next call or at check_errors()."""
    global _CHECK
    if mode not in ("immediate", "deferred"):
        raise ValueError(mode)
    _CHECK = mode


_POOL = collections.OrderedDict()     # call signature -> prepared calls (ops.KldCall / SplitKldCall / LatentTailKldCall)
_POOL_SIGNATURES, _POOL_DEPTH = 8, 4


def _pooled_call(key, make):
    """A prepared call for this problem signature, reused across training steps: building one allocates the workspace (GBs
    at M = 256) and a dozen output buffers, and with several live at once (deferred error checks, autograd graphs) the
    caching allocator kept going back to cudaMalloc — 16 .. 63 cudaMalloc calls inside five timed steps on two GPUs, where
    every cudaMalloc also has to map the block into the peers.  A call is free when NOTHING refers to it any more (the
    autograd node of its forward, the deferred-check queue, the H^-1 / latent-tail tags on grad_H): plain reference
    counting, so a caller that keeps two graphs alive simply gets two calls.  Work is stream-ordered: the stream is part of
    the key."""
    calls = _POOL.get(key)
    if calls is None:
        calls = _POOL[key] = []
        while len(_POOL) > _POOL_SIGNATURES:          # ragged minibatches: every new shape is a new signature
            _POOL.popitem(last=False)
    else:
        _POOL.move_to_end(key)
    for c in calls:
        if sys.getrefcount(c) <= 3:                   # the pool's list, this loop variable, getrefcount's argument
            return c
    c = make()
    if len(calls) < _POOL_DEPTH:
        calls.append(c)
    return c


def clear_call_pool():
    """Drop the cached calls (their device memory returns to torch's allocator)."""
    _POOL.clear()


def check_errors():
    """Raise if any earlier deferred call hit a non-positive-definite block."""
    while _PENDING:
        _PENDING.pop(0).raise_on_info()


def _noise_of(likelihood, L, dtype, device):
    src = getattr(likelihood, "noise_covar", likelihood)
    return src.noise.reshape(-1).to(dtype=dtype, device=device).expand(L)


def _noise_entry(likelihood):
    """The likelihood noise as a lazy Raw reference when its transform is known (packed evaluation with the kernel
    hyper-parameters, spec.build_structure), else the evaluated tensor."""
    from .constraints import GreaterThan, Positive
    from .spec import Raw
    src = getattr(likelihood, "noise_covar", likelihood)
    con = getattr(src, "raw_noise_constraint", None)
    if hasattr(src, "raw_noise") and type(con) in (GreaterThan, Positive):
        return Raw(src.raw_noise, "softplus", con.lower_float())
    if hasattr(src, "_log_noise") and hasattr(src, "min_log_noise"):
        from .GP_model import _min_float
        return Raw(src._log_noise, "bounded", _min_float(src, "min_log_noise"))
    return src.noise.reshape(-1)


class _KldBound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, log_v, m, H, hyper, meta):
        """hyper: [n_ls + n_comp + 1, L] = lengthscales | outputscales | noise (constrained values)."""
        st = meta["structure"]
        lengthscale, outputscale, noise = hyper[:st.n_ls], hyper[st.n_ls:st.n_ls + st.n_comp], hyper[st.n_ls + st.n_comp]
        x, z, offsets = meta["x"], meta["z"], meta["offsets"]
        L, M, Q = meta["L"], H.shape[-1], x.shape[1]
        counts = meta.get("counts")
        latent_tail = meta.get("group") is not None and meta.get("tail") == "latents"

        def make():
            if counts is not None:                    # ragged minibatch: rows per subject known on the host
                c = ops.make_kld_call(st, L, M, Q, counts, x.device, natural_gradient=meta["natural_gradient"], path=_PATH)
            else:
                c = ops.KldCall(st, L, M, Q, offsets.numel() - 1, x.shape[0], meta["T_max"], meta["sum_T2"], x.device,
                                natural_gradient=meta["natural_gradient"], path=_PATH)
            if latent_tail and isinstance(c, ops.KldCall):
                c = ops.LatentTailKldCall(c, meta["group"])
            return c
        key = (st.table.tobytes(), st.n_comp0, st.n_comp1, st.n_ls, L, M, Q, offsets.numel() - 1, x.shape[0], meta["T_max"],
               meta["sum_T2"], x.device, meta["natural_gradient"], _PATH, id(meta["group"]) if latent_tail else 0,
               None if counts is None else np.asarray(counts, dtype=np.int64).tobytes(),
               torch.cuda.current_stream(x.device).cuda_stream)
        call = _pooled_call(key, make)
        call.bind(x, offsets, mu, log_v, z, m.reshape(L, M), H, lengthscale, outputscale, noise, meta["scale"],
                  meta["const_term"], meta["eps"])
        call.head()
        exchange_stats(call, meta.get("group"), _EXCHANGE)
        call.tail()
        if _CHECK == "immediate":
            call.raise_on_info()
        else:
            call.post_info()
            _PENDING.append(call)
            if len(_PENDING) > 2:                      # the oldest one finished long ago: reading its flags does not stall
                _PENDING.pop(0).raise_on_info()
        kld = call.kld_per_latent.sum()
        ctx.call = call
        ctx.ng = meta["natural_gradient"]
        ctx.mshape = m.shape
        # the call object (scratch + outputs) goes back to the pool once nothing refers to it: hand out copies of the two
        # small results that outlive it
        gm, gH = call.grad_m.view(L, M, 1).clone(), call.grad_H.clone()
        ctx.mark_non_differentiable(gm, gH)
        # H^-1 of the head kernel rides along for natural_gradient_step (training.py:130-131 recomputes it); the tag keeps
        # the call (whose workspace holds it) out of the pool for as long as grad_H lives
        meta["Hinv"] = (call.Hinv, H.data_ptr(), H._version, call)
        meta["latent_tail"] = call if isinstance(call, ops.LatentTailKldCall) else None
        return kld, gm, gH

    @staticmethod
    def backward(ctx, g, _gm, _gH):
        c = ctx.call
        d_m = d_H = None
        if not ctx.ng:
            d_m, d_H = g * c.grad_m.view(ctx.mshape), g * c.grad_H
        d_mu, d_lv, d_hyp = torch._foreach_mul([c.d_mu, c.d_log_v, c.d_hyper], g)       # one grouped launch
        return (d_mu, d_lv, d_m, d_H, d_hyp, None)


def _structure_of(covar_module0, covar_module1, L, device):
    return build_structure(flatten(covar_module0), flatten(covar_module1), L, device=device)


def _run(covar_module0, covar_module1, likelihood, latent_dim, m, H, x, offsets, T_max, sum_T2, mu, log_v, z, scale,
         const_term, natural_gradient, eps, counts=None):
    _need_cuda(x)
    L = latent_dim
    f64 = torch.float64
    st, ls, os_, nz = build_structure(flatten(covar_module0), flatten(covar_module1), L, device=x.device,
                                      extra=[_noise_entry(likelihood)])
    # one [n_ls + n_comp + 1, L] table; when the packed transform produced it, the three parts are views of one tensor
    same = ls._base is not None and ls._base is os_._base and ls._base is nz._base and ls._base.shape[0] == st.n_ls + st.n_comp + 1
    hyper = ls._base if same else torch.cat([ls, os_, nz.reshape(1, L).to(f64)])
    if z.dim() == 2:
        z = z.unsqueeze(0).expand(L, -1, -1)
    if int(T_max) > MAX_T:                                   # longer subjects than the fused kernels cover
        if _GROUP is not None:
            raise RuntimeError(f"lvae_b200: subjects with more than {MAX_T} rows are not supported in the sharded modes")
        if counts is None:
            counts = np.full(offsets.numel() - 1, int(T_max), dtype=np.int64)
        return _composed_bound(st, ls, os_, nz.reshape(L), L, m, H, x, offsets, counts, mu, log_v, z, scale, const_term,
                               natural_gradient, eps)
    meta = dict(structure=st, x=x.to(f64), z=z.to(f64), offsets=offsets, L=L, T_max=int(T_max), sum_T2=int(sum_T2),
                scale=scale, const_term=const_term, eps=eps, natural_gradient=bool(natural_gradient), counts=counts)
    if _GROUP is not None and _SHARD == "latents":
        return _run_latent_shard(mu.to(f64), log_v.to(f64), m.to(f64), H.to(f64), hyper, meta, natural_gradient)
    meta["group"] = _GROUP
    meta["tail"] = _TAIL if (_GROUP is not None and natural_gradient) else "replicated"
    kld, gm, gH = _KldBound.apply(mu.to(f64), log_v.to(f64), m.to(f64), H.to(f64), hyper, meta)
    if natural_gradient:
        gH._lvae_hinv = meta.get("Hinv")
        gH._lvae_latent_tail = meta.get("latent_tail")     # grad_m / grad_H hold this rank's latents only (see ops.LatentTailKldCall)
        return kld, gm, gH
    return kld, None, None


def _composed_bound(st, ls, os_, noise, L, m, H, x, offsets, counts, mu, log_v, z, scale, const_term, natural_gradient, eps):
    """The same bound for subjects with MORE than 40 rows (up to 256), which the fused kernels do not cover: composed from the
    differentiable CUDA ops of diff_ops.py (kernel matrices, batched SPD inverse / log-det, DMMA GEMM), autograd instead of
    the closed-form adjoints.  Formulas: SURVEY 8(a) / elbo_functions.py:171-214, 264-305; subjects are processed in groups of
    equal length (one batched factorisation per distinct T).  An order of magnitude slower per row than the fused path —
    a correctness fallback, not a throughput path."""
    from . import diff_ops as D
    f64 = torch.float64
    dev = x.device
    x = x.detach().to(f64).contiguous()
    z = z.detach().to(f64).contiguous()
    N, M = x.shape[0], z.shape[1]
    counts = np.asarray(counts, dtype=np.int64)
    if int(counts.max()) > MAX_M:
        raise RuntimeError(f"lvae_b200: a subject has {int(counts.max())} rows; the batched Cholesky covers up to {MAX_M}")
    row0 = np.concatenate([[0], np.cumsum(counts)])
    blk0 = np.concatenate([[0], np.cumsum(counts * counts)])
    sum_T2 = int(blk0[-1])
    muT, lvT = mu.to(f64).t().contiguous(), log_v.to(f64).t().contiguous()               # [L, N]
    m3, H = m.to(f64).reshape(L, M, 1), H.to(f64)
    noise = noise.to(f64).contiguous()
    Kxz = D.KernelDense.apply(st, "k0", x, z, ls, os_, None)                              # [L, N, M]      (171)
    Kzz = D.KernelDense.apply(st, "k0", z, z, ls, os_, None) + eps * torch.eye(M, dtype=f64, device=dev)   # (172, 176)
    K0b = D.KernelBlocks.apply(st, "k0", x, offsets, sum_T2, ls, os_, None)               # flat [L, sum T^2]  (173)
    Bb = D.KernelBlocks.apply(st, "k1", x, offsets, sum_T2, ls, os_, noise)               # K1 + noise I       (174)
    Ki, ldK = D.spd_inverse(Kzz)                                                          # (177-178)
    Hi, ldH = D.spd_inverse(H)                                                            # (185-186)
    a = D.gemm(Ki, m3)
    r = D.gemm(Kxz, a).reshape(L, N) - muT                                                # (189)
    zero = torch.zeros(L, dtype=f64, device=dev)
    S, ng1 = torch.zeros(L, M, M, dtype=f64, device=dev), torch.zeros(L, M, 1, dtype=f64, device=dev)
    A, Bt, C, D1 = zero, zero, zero, zero
    for T in np.unique(counts).tolist():
        sel = np.nonzero(counts == T)[0]
        PT = len(sel)
        rows = torch.from_numpy((row0[sel][:, None] + np.arange(T)[None, :]).reshape(-1)).to(dev)
        ent = torch.from_numpy((blk0[sel][:, None] + np.arange(T * T)[None, :]).reshape(-1)).to(dev)
        iB, ldB = D.spd_inverse(Bb[:, ent].reshape(L * PT, T, T))                         # (179-180)
        Kx = Kxz[:, rows].reshape(L * PT, T, M)
        V = D.gemm(iB, Kx)                                                                # B^-1 Kxz
        S = S + D.gemm(Kx.reshape(L, PT * T, M), V.reshape(L, PT * T, M), ta=True)        # (183-184)
        rt = r[:, rows].reshape(L * PT, T, 1)
        A = A + (rt * D.gemm(iB, rt)).reshape(L, -1).sum(1)                               # (190)
        Bt = Bt + (torch.diagonal(iB, dim1=-2, dim2=-1).reshape(L, -1) * torch.exp(lvT[:, rows])).sum(1)       # (191)
        C = C + ldB.reshape(L, PT).sum(1)                                                 # (192)
        D1 = D1 + (iB * K0b[:, ent].reshape(L * PT, T, T)).reshape(L, -1).sum(1)          # (193), first term
        ng1 = ng1 + D.gemm(V.reshape(L, PT * T, M), muT[:, rows].reshape(L, PT * T, 1), ta=True)               # (208)
    Dt = D1 - (S * Ki).reshape(L, -1).sum(1)                                              # (193)
    G = D.gemm(D.gemm(Ki, H), Ki)                                                         # (194)
    E = (G * S).reshape(L, -1).sum(1)                                                     # (195 / 282)
    F_ = lvT.sum(1)                                                                       # (196)
    kl_u = 0.5 * ((Ki * H.transpose(1, 2)).reshape(L, -1).sum(1) + (m3 * a).reshape(L, -1).sum(1) - M + ldK - ldH)   # (199-203)
    kld = (scale * 0.5 * (A + Bt + C + Dt + E - F_) + kl_u).sum() - const_term            # (204 / 299)
    if not natural_gradient:
        return kld, None, None
    with torch.no_grad():                                                                 # (208-214 / 301-305)
        Bm = D.gemm(D.gemm(Ki, S), Ki) + Ki
        grad_m = D.gemm(Bm, m3) - D.gemm(Ki, ng1)
        grad_H = 0.5 * (Bm - Hi)
    return kld, grad_m, grad_H


def latent_slice(L, rank, world):
    """[l0, l1) of `rank` when L latent dimensions are dealt to `world` ranks in contiguous, balanced blocks."""
    return (L * rank) // world, (L * (rank + 1)) // world


def _run_latent_shard(mu, log_v, m, H, hyper, meta, natural_gradient):
    """Latent-dimension sharding (the fallback for minibatches too small to split by subject, where the per-latent
    M x M work dominates).  The latent dimensions of the bound are independent, so rank r runs the whole op on latents
    [l0, l1) of the SAME minibatch: kld_total is all-reduced (its gradient reaches this rank's latent columns of mu, log_v and
    of every hyper-parameter; summing the parameter gradients over ranks gives the full gradient), grad_m / grad_H are
    all-gathered into full [L, ...] tensors."""
    import torch.distributed as dist
    L, M = meta["L"], H.shape[-1]
    rank, world = dist.get_rank(_GROUP), dist.get_world_size(_GROUP)
    if L < world:
        raise RuntimeError(f"lvae_b200: latent sharding needs at least one latent dimension per rank (L={L}, ranks={world})")
    l0, l1 = latent_slice(L, rank, world)
    Ll = l1 - l0
    sub = dict(meta)
    sub.update(L=Ll, z=meta["z"][l0:l1].contiguous(), const_term=meta["const_term"] * Ll / L, group=None)
    kld_l, gm_l, gH_l = _KldBound.apply(mu[:, l0:l1].contiguous(), log_v[:, l0:l1].contiguous(),
                                        m.reshape(L, M, 1)[l0:l1].contiguous(), H[l0:l1].contiguous(),
                                        hyper[:, l0:l1].contiguous(), sub)
    tot = kld_l.detach().clone()
    dist.all_reduce(tot, group=_GROUP)
    kld = kld_l + (tot - kld_l.detach())                    # value: all latents; gradient: this rank's latents
    if not natural_gradient:
        return kld, None, None
    Lmax = (L + world - 1) // world                          # blocks differ by at most one latent: pad to the largest
    send = torch.zeros(Lmax, M * (M + 1), dtype=torch.float64, device=mu.device)
    send[:Ll, :M] = gm_l.reshape(Ll, M)
    send[:Ll, M:] = gH_l.reshape(Ll, M * M)
    recv = torch.empty(world, Lmax, M * (M + 1), dtype=torch.float64, device=mu.device)
    dist.all_gather_into_tensor(recv, send, group=_GROUP)
    parts = [recv[r, :latent_slice(L, r, world)[1] - latent_slice(L, r, world)[0]] for r in range(world)]
    full = torch.cat(parts)
    return kld, full[:, :M].reshape(L, M, 1).contiguous(), full[:, M:].reshape(L, M, M).contiguous()


def minibatch_KLD_upper_bound(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, mu, log_v, z, P_tot,
                              P_batch, T, natural_gradient, eps):
    """Unbiased minibatch estimate of the KL upper bound, fixed T rows per subject (elbo_functions.py:144-216).

    train_xt [N_b,Q] is subject-major with exactly T rows per subject (the reference reshapes without checking ids,
    line 168).  Returns (kld_total, grad_m [L,M,1], grad_H [L,M,M]); the natural-gradient terms are not scaled by
    P_tot/P_batch, as in the reference (208-214)."""
    N_b = train_xt.shape[0]
    P_loc = P_batch if _GROUP is None else N_b // T
    if P_loc * T != N_b:
        raise RuntimeError(f"shape '[{P_batch}, {T}, {train_xt.shape[1]}]' is invalid for input of size {train_xt.numel()}")
    offsets = torch.arange(0, N_b + 1, T, dtype=torch.int32, device=train_xt.device)
    return _run(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, offsets, T, P_loc * T * T, mu,
                log_v, z, P_tot / P_batch, latent_dim * P_tot * T / 2, natural_gradient, eps)


def group_by_subject(ids):
    """Subject grouping of elbo_functions.py:264-267, bit-exact: subjects = sorted unique ids; rows of a subject in their
    original order.  Returns (row order or None if already grouped, offsets int32 [P+1] on the device, T_max, sum_T2)."""
    uniq, inverse, counts = torch.unique(ids, sorted=True, return_inverse=True, return_counts=True)
    order = torch.argsort(inverse, stable=True)
    counts_h = counts.cpu()                       # one sync; torch.unique(...).tolist() in the reference syncs as well
    offsets = torch.zeros(counts_h.numel() + 1, dtype=torch.int64)
    torch.cumsum(counts_h, 0, out=offsets[1:])
    grouped = bool((order == torch.arange(order.numel(), device=order.device)).all())
    group_by_subject.last_counts = counts_h.numpy()
    return (None if grouped else order, offsets.to(torch.int32).to(ids.device), int(counts_h.max()),
            int((counts_h * counts_h).sum()))


def subject_counts_of(id_column):
    """Rows per subject when the rows of every subject are CONTIGUOUS (what the reference's samplers produce, utils.py:79-113),
    else None.  `id_column`: host tensor / array of the id covariate of a minibatch — call it on the loader's CPU batch, before
    the copy to the device, and pass the result as `subject_counts=`: the bound then needs no torch.unique and no
    device->host synchronisation to find its subjects."""
    ids = torch.as_tensor(id_column).reshape(-1).cpu()
    if ids.numel() == 0:
        return np.zeros(0, dtype=np.int64)
    uniq, counts = torch.unique_consecutive(ids, return_counts=True)
    if torch.unique(uniq).numel() != uniq.numel():          # an id comes back after another subject: rows are not grouped
        return None
    return counts.numpy().astype(np.int64)


def minibatch_KLD_upper_bound_iter(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, mu, log_v, z, P,
                                   P_in_current_batch, N, natural_gradient, id_covariate, eps, subject_counts=None):
    """Same bound for irregular numbers of rows per subject (elbo_functions.py:219-307): subjects are the sorted unique
    values of column `id_covariate`; the constant term is L*N/2 with N the number of rows of the whole data set.
    subject_counts (optional, not in the reference): host array of rows per subject for a minibatch whose subjects occupy
    contiguous row blocks (see subject_counts_of) — the grouping of lines 264-267 is then known without looking at the ids on
    the device; the bound is a sum over subjects, so their order does not matter."""
    if subject_counts is not None:
        counts = np.asarray(subject_counts, dtype=np.int64)
        if int(counts.sum()) != train_xt.shape[0]:
            raise RuntimeError("lvae_b200: subject_counts does not add up to the number of rows of the minibatch")
        off = np.zeros(counts.size + 1, dtype=np.int32)
        np.cumsum(counts, out=off[1:])
        offsets = torch.from_numpy(off).to(train_xt.device, non_blocking=True)
        T_max, sum_T2 = int(counts.max()) if counts.size else 0, int((counts * counts).sum())
    else:
        order, offsets, T_max, sum_T2 = group_by_subject(train_xt[:, id_covariate])
        if order is not None:
            train_xt, mu, log_v = train_xt[order], mu[order], log_v[order]
        counts = group_by_subject.last_counts
    kld, gm, gH = _run(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, offsets, T_max, sum_T2, mu,
                       log_v, z, P / P_in_current_batch, latent_dim * N / 2, natural_gradient, eps, counts=counts)
    return kld.reshape(1), gm, gH


# ---------------------------------------------------------------------------------------------------------------
# the non-minibatch bounds of elbo_functions.py:8-142 and validation.py:8-68.  Like the reference's, the values
# are autograd graph nodes w.r.t. the variational mean / log-variance (or the latent sample), every kernel hyper-parameter
# and the likelihood noise, so the loops that minimise them (training.py:326-343, 533-548, 654, 730) work unchanged; the
# covariates and inducing inputs are constants (LVAE.py:204-208).  Kernel matrices and their hyper-parameter adjoints come
# from lvae_kernel_dense/blocks[_bwd]_f64, inverses and log-determinants from lvae_potrf/potri_batched_f64, the
# contractions from the batched DMMA GEMM (diff_ops.py); the element-wise glue between them is plain torch.
# ---------------------------------------------------------------------------------------------------------------
def _need_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("lvae_b200: the GP-prior bounds need CUDA tensors (no CPU fallback)")


def _stack_latents(modules0, modules1, likelihoods, device):
    """(structure, lengthscale [n_ls,L], outputscale [n_comp,L], noise [L]) from L UN-batched kernel modules of identical
    structure (the per-latent lists `covar_module0[i]`, `covar_module1[i]`, `likelihoods[i]` of the non-Hensman loops,
    LVAE.py:244-270): latent l's hyper-parameters become column l of the tables, differentiably."""
    from .spec import FlatComponent, Raw, _val
    L = len(modules0)
    if not (L == len(modules1) == len(likelihoods)) or L == 0:
        raise RuntimeError("lvae_b200: need one (covar_module0, covar_module1, likelihood) triple per latent dimension")

    def col(t):
        t = _val(t).reshape(-1)
        if t.numel() != 1:
            raise RuntimeError("lvae_b200: the per-latent modules must be un-batched (one value per hyper-parameter)")
        return t

    def stack(entries):
        """One [L] entry from the L per-latent ones.  Lazy entries (raw parameter + constraint) stay lazy: their raw values
        are concatenated and build_structure applies ONE transform to the whole table instead of one per module."""
        if all(isinstance(e, Raw) and e.param.numel() == 1 for e in entries) and \
                len({(e.kind, e.lower) for e in entries}) == 1:
            return Raw(torch.cat([e.param.reshape(1) for e in entries]), entries[0].kind, entries[0].lower)
        return torch.cat([col(e) for e in entries])

    def merge(per_latent):
        out = []
        for ci, first in enumerate(per_latent[0]):
            shape = [(k, d) for k, d, _ in first.factors]
            cs = [pl[ci] for pl in per_latent]
            if any(len(pl) != len(per_latent[0]) for pl in per_latent) or \
                    any([(k, d) for k, d, _ in c.factors] != shape or (c.outputscale is None) != (first.outputscale is None)
                        for c in cs):
                raise RuntimeError("lvae_b200: the per-latent kernel modules must share one structure")
            os_ = None if first.outputscale is None else stack([c.outputscale for c in cs])
            factors = [(k, d, None if ls is None else stack([c.factors[fi][2] for c in cs]))
                       for fi, (k, d, ls) in enumerate(first.factors)]
            out.append(FlatComponent(os_, factors))
        return out

    st, ls, os_, nz = build_structure(merge([flatten(m) for m in modules0]), merge([flatten(m) for m in modules1]), L,
                                      device=device, extra=[stack([_noise_entry(lk) for lk in likelihoods])])
    return st, ls, os_, nz.reshape(L).to(torch.float64)


def _low_rank_terms(L, covar_module0, covar_module1, likelihood, x, z, P, T, eps, hyper=None):
    """Shared pieces of elbo / deviance_upper_bound, per latent l (FP64, on x's device):
    K0xz [L,N,M], iB [L*P,T,T], iB_K0xz [L,N,M], S = K0zx iB K0xz, iW = (Kzz + S)^-1, the summed log-dets, tr.
    hyper: precomputed (structure, lengthscale, outputscale, noise) instead of the three modules."""
    from . import diff_ops as D
    _need_cuda(x)
    f64 = torch.float64
    x = x.detach().to(f64).contiguous()
    z = z.detach().to(f64)
    if z.dim() == 2:
        z = z.unsqueeze(0).expand(L, -1, -1)
    z = z.contiguous()
    N, M = x.shape[0], z.shape[1]
    if N != P * T:
        raise RuntimeError(f"shape '[{P}, {T}, {x.shape[1]}]' is invalid for input of size {x.numel()}")
    if hyper is not None:
        st, ls, os_, noise = hyper
        noise = noise.contiguous()
    else:
        st, ls, os_ = _structure_of(covar_module0, covar_module1, L, x.device)
        noise = _noise_of(likelihood, L, f64, x.device).contiguous()
    offsets = torch.arange(0, N + 1, T, dtype=torch.int32, device=x.device)
    eye = torch.eye(M, dtype=f64, device=x.device)
    K0xz = D.KernelDense.apply(st, "k0", x, z, ls, os_, None)
    K0zz = D.KernelDense.apply(st, "k0", z, z, ls, os_, None) + eps * eye
    K0_st = D.KernelBlocks.apply(st, "k0", x, offsets, P * T * T, ls, os_, None).reshape(L * P, T, T)
    B_st = D.KernelBlocks.apply(st, "k1", x, offsets, P * T * T, ls, os_, noise).reshape(L * P, T, T)
    iK, ldK = D.spd_inverse(K0zz)
    iB, ldB = D.spd_inverse(B_st)
    iB_K0xz = D.gemm(iB, K0xz.reshape(L * P, T, M)).reshape(L, N, M)
    S = D.gemm(K0xz, iB_K0xz, ta=True)
    W = K0zz + S
    W = 0.5 * (W + W.transpose(1, 2))
    iW, ldW = D.spd_inverse(W)
    logDet = -ldK + ldB.reshape(L, P).sum(1) + ldW
    tr = (iB * K0_st).reshape(L, -1).sum(1) - (S * iK).reshape(L, -1).sum(1)
    return dict(K0xz=K0xz, iB=iB, iB_K0xz=iB_K0xz, iW=iW, logDet=logDet, tr=tr, N=N, M=M)


def _quad_form(t, y, L, P, T):
    """qF = y^T B^-1 y - p^T W^-1 p with p = K0zx B^-1 y, per latent; y [L,N]."""
    from . import diff_ops as D
    iB_y = D.gemm(t["iB"], y.reshape(L * P, T, 1)).reshape(L, -1)
    qF1 = (y * iB_y).sum(1)
    p = D.gemm(t["K0xz"], iB_y.unsqueeze(2), ta=True)
    qF2 = (p * D.gemm(t["iW"], p)).reshape(L, -1).sum(1)
    return qF1 - qF2


def _dubo_per_latent(L, covar_module0, covar_module1, likelihood, train_xt, m, log_v, z, P, T, eps, hyper=None):
    """elbo_functions.py:90-142 / validation.py:8-68 for all latents at once; m, log_v [N,L] (or [N] for L = 1)."""
    from . import diff_ops as D
    t = _low_rank_terms(L, covar_module0, covar_module1, likelihood, train_xt, z, P, T, eps, hyper)
    f64 = torch.float64
    mL = m.to(f64).reshape(t["N"], L).t().contiguous()
    lv = log_v.to(f64).reshape(t["N"], L).t().contiguous()
    v = torch.exp(lv)
    qF = _quad_form(t, mL, L, P, T)
    tr_iB_D = (torch.diagonal(t["iB"], dim1=-2, dim2=-1).reshape(L, -1) * v).sum(1)
    D05 = t["iB_K0xz"] * torch.sqrt(v).unsqueeze(2)
    SD = D.gemm(D05, D05, ta=True, flags=3)                                      # K0zx B^-1 D B^-1 K0xz (symmetric)
    tr_iSigma_D = tr_iB_D - (t["iW"] * SD).reshape(L, -1).sum(1)
    return 0.5 * (tr_iSigma_D + qF - P * T + t["logDet"] - lv.sum(1) + t["tr"])


def deviance_upper_bound(covar_module0, covar_module1, likelihood, train_xt, m, log_v, z, P, T, eps):
    """DUBO of one latent dimension with un-batched kernels (elbo_functions.py:90-142); differentiable (see above)."""
    return _dubo_per_latent(1, covar_module0, covar_module1, likelihood, train_xt, m, log_v, z, P, T, eps).reshape(())


def _elbo_per_latent(L, covar_module0, covar_module1, likelihood, train_xt, train_yt, z, P, T, eps, hyper=None):
    import math
    t = _low_rank_terms(L, covar_module0, covar_module1, likelihood, train_xt, z, P, T, eps, hyper)
    y = train_yt.to(torch.float64).reshape(t["N"], L).t().contiguous()
    qF = _quad_form(t, y, L, P, T)
    logLike = -0.5 * T * P * math.log(2 * math.pi) - 0.5 * (t["logDet"] + qF)
    return logLike - 0.5 * t["tr"]


def elbo(covar_module0, covar_module1, likelihood, train_xt, train_yt, z, P, T, eps):
    """Low-rank evidence lower bound of one latent dimension given a latent sample (elbo_functions.py:36-88);
    differentiable w.r.t. the sample, the kernel hyper-parameters and the noise."""
    return _elbo_per_latent(1, covar_module0, covar_module1, likelihood, train_xt, train_yt, z, P, T, eps).reshape(())


def _z_stack(zt_list):
    return zt_list if torch.is_tensor(zt_list) else torch.stack([z for z in zt_list])


def deviance_upper_bound_all(covar_modules0, covar_modules1, likelihoods, train_xt, m, log_v, zt_list, P, T, eps):
    """The per-latent loop of training.py:334-343 / 537-546 / 654 / 730 as ONE batched evaluation (not in the reference):
    `covar_modules0[i]`, `covar_modules1[i]`, `likelihoods[i]`, `zt_list[i]` are the un-batched per-latent objects of the
    non-Hensman path, m / log_v are [N, L].  Returns the [L] vector whose entry i equals
    deviance_upper_bound(covar_modules0[i], covar_modules1[i], likelihoods[i], train_xt, m[:, i], log_v[:, i], zt_list[i], ...);
    `.sum()` is the reference's accumulated gp_loss.  Differentiable like the single-latent function."""
    L = len(covar_modules0)
    hyper = _stack_latents(covar_modules0, covar_modules1, likelihoods, train_xt.device)
    return _dubo_per_latent(L, None, None, None, train_xt, m, log_v, _z_stack(zt_list), P, T, eps, hyper)


def elbo_all(covar_modules0, covar_modules1, likelihoods, train_xt, Z, zt_list, P, T, eps):
    """The per-latent loop of training.py:324-329 / 529-535 as one batched evaluation: entry i of the returned [L] vector is
    elbo(covar_modules0[i], covar_modules1[i], likelihoods[i], train_xt, Z[:, i], zt_list[i], P, T, eps)."""
    L = len(covar_modules0)
    hyper = _stack_latents(covar_modules0, covar_modules1, likelihoods, train_xt.device)
    return _elbo_per_latent(L, None, None, None, train_xt, Z, _z_stack(zt_list), P, T, eps, hyper)


def KL_closed(covar_module, train_x, likelihoods, data, mu, log_var):
    """Closed-form KL[q || GP prior] with the dense N x N kernel (elbo_functions.py:8-34); N <= 256 (the batched Cholesky's
    limit — the reference uses this for small exact checks only).  Differentiable like the two bounds above."""
    from . import diff_ops as D
    f64 = torch.float64
    _need_cuda(train_x)
    N = data.shape[0]
    x = train_x.detach().to(f64)
    noise = _noise_of(likelihoods, 1, f64, x.device)
    K1 = covar_module(x, x).evaluate().reshape(N, N) + noise * torch.eye(N, dtype=f64, device=x.device)
    iK, logdet11 = D.spd_inverse(K1.unsqueeze(0))
    iK = iK[0]
    mu1 = mu.to(f64).reshape(-1)
    lv1 = log_var.to(f64).reshape(-1)
    qf1 = (mu1 * (iK @ mu1)).sum()
    tr1 = (torch.exp(lv1) * torch.diagonal(iK)).sum()
    return 0.5 * (tr1 + qf1 - N + logdet11.sum() - lv1.sum())
