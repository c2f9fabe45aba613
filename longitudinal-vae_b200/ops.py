"""Thin Python wrappers over the C ABI (include/lvae_b200.h): tensor allocation, pointer passing, error mapping."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import KldProblemT, check, make_spec, ptr, require_cuda, stream_ptr

F64 = torch.float64


def _c(t):
    t = t.detach()
    if t.dtype is F64 and t.is_contiguous():         # the common case: nothing to convert (this runs ~20 times per step)
        return t
    return t.to(F64).contiguous()


def kernel_dense(structure, lengthscale, outputscale, x1, x2, which="all", diag_add=None):
    """Dense additive kernel [B, n1, n2] on the GPU.  x: [n,Q] (shared) or [B,n,Q]; B a multiple of L (matrix b uses the
    hyper-parameters of latent b % L).  which: "k0" | "k1" | "all".  Replaces covar_module(x1,x2).evaluate()."""
    lib = require_cuda(x1, x2, lengthscale, outputscale)
    L = outputscale.shape[1]
    lo, hi = {"k0": (0, structure.n_comp0), "k1": (structure.n_comp0, structure.n_comp),
              "all": (0, structure.n_comp)}[which]
    x1, x2 = _c(x1), _c(x2)
    B = max(L, x1.shape[0] if x1.dim() == 3 else 1, x2.shape[0] if x2.dim() == 3 else 1)
    s1 = x1.shape[-2] * x1.shape[-1] if x1.dim() == 3 else 0
    s2 = x2.shape[-2] * x2.shape[-1] if x2.dim() == 3 else 0
    for t in (x1, x2):
        if t.dim() == 3 and t.shape[0] != B:
            raise RuntimeError("lvae_b200: batched covariates must share the leading dimension")
    n1, n2, Q = x1.shape[-2], x2.shape[-2], x1.shape[-1]
    out = torch.empty(B, n1, n2, dtype=F64, device=x1.device)
    ks, keep = make_spec(structure)
    ls, os_ = _c(lengthscale), _c(outputscale)
    da = _c(diag_add) if diag_add is not None else None
    with torch.cuda.device(x1.device):
        rc = lib.lvae_kernel_dense_f64(C.byref(ks), lo, hi, L, B, Q, ptr(x1), s1, n1, ptr(x2), s2, n2, ptr(ls),
                                       ptr(os_), ptr(da), ptr(out), stream_ptr(x1.device))
    check(rc, "lvae_kernel_dense_f64")
    return out


def kernel_blocks(structure, lengthscale, outputscale, x, offsets_dev, sum_T2, which="k0", diag_add=None):
    """Per-subject blocks, flat [L, sum_T2] (block p of latent l at offset sum_{q<p} T_q^2, row-major T_p x T_p)."""
    lib = require_cuda(x, lengthscale, outputscale, offsets_dev)
    L = outputscale.shape[1]
    lo, hi = {"k0": (0, structure.n_comp0), "k1": (structure.n_comp0, structure.n_comp),
              "all": (0, structure.n_comp)}[which]
    x = _c(x)
    P_b = offsets_dev.numel() - 1
    out = torch.empty(L, int(sum_T2), dtype=F64, device=x.device)
    ks, keep = make_spec(structure)
    ls, os_ = _c(lengthscale), _c(outputscale)
    da = _c(diag_add) if diag_add is not None else None
    with torch.cuda.device(x.device):
        rc = lib.lvae_kernel_blocks_f64(C.byref(ks), lo, hi, L, x.shape[1], ptr(x), ptr(offsets_dev), P_b, int(sum_T2),
                                        ptr(ls), ptr(os_), ptr(da), ptr(out), stream_ptr(x.device))
    check(rc, "lvae_kernel_blocks_f64")
    return out


def _comp_range(structure, which):
    return {"k0": (0, structure.n_comp0), "k1": (structure.n_comp0, structure.n_comp), "all": (0, structure.n_comp)}[which]


def kernel_dense_bwd(structure, lengthscale, outputscale, x1, x2, grad_out, which="all", want_diag=False):
    """Adjoints of kernel_dense w.r.t. (lengthscale [n_ls,L], outputscale [n_comp,L], diag_add [L] | None) given the adjoint
    of its output (lvae_kernel_dense_bwd_f64)."""
    lib = require_cuda(x1, x2, lengthscale, outputscale, grad_out)
    L = outputscale.shape[1]
    lo, hi = _comp_range(structure, which)
    x1, x2, g = _c(x1), _c(x2), _c(grad_out)
    B = g.shape[0]
    s1 = x1.shape[-2] * x1.shape[-1] if x1.dim() == 3 else 0
    s2 = x2.shape[-2] * x2.shape[-1] if x2.dim() == 3 else 0
    n1, n2, Q = x1.shape[-2], x2.shape[-2], x1.shape[-1]
    if g.shape != (B, n1, n2) or B % L:
        raise RuntimeError("lvae_b200: kernel_dense_bwd: adjoint shape does not match the kernel matrix")
    ls, os_ = _c(lengthscale), _c(outputscale)
    d_ls, d_os = torch.empty_like(ls), torch.empty_like(os_)
    d_diag = torch.empty(L, dtype=F64, device=g.device) if want_diag else None
    ks, keep = make_spec(structure)
    with torch.cuda.device(g.device):
        rc = lib.lvae_kernel_dense_bwd_f64(C.byref(ks), lo, hi, L, B, Q, ptr(x1), s1, n1, ptr(x2), s2, n2, ptr(ls), ptr(os_),
                                           ptr(g), ptr(d_ls), ptr(d_os), ptr(d_diag), stream_ptr(g.device))
    check(rc, "lvae_kernel_dense_bwd_f64")
    return d_ls, d_os, d_diag


def kernel_blocks_bwd(structure, lengthscale, outputscale, x, offsets_dev, grad_out, which="k0", want_diag=False):
    """Adjoints of kernel_blocks (lvae_kernel_blocks_bwd_f64); grad_out flat [L, sum_T2]."""
    lib = require_cuda(x, lengthscale, outputscale, offsets_dev, grad_out)
    L = outputscale.shape[1]
    lo, hi = _comp_range(structure, which)
    x, g = _c(x), _c(grad_out)
    P_b = offsets_dev.numel() - 1
    if g.dim() != 2 or g.shape[0] != L:
        raise RuntimeError("lvae_b200: kernel_blocks_bwd: adjoint must be [L, sum_T2]")
    ls, os_ = _c(lengthscale), _c(outputscale)
    d_ls, d_os = torch.empty_like(ls), torch.empty_like(os_)
    d_diag = torch.empty(L, dtype=F64, device=g.device) if want_diag else None
    ks, keep = make_spec(structure)
    with torch.cuda.device(g.device):
        rc = lib.lvae_kernel_blocks_bwd_f64(C.byref(ks), lo, hi, L, x.shape[1], ptr(x), ptr(offsets_dev), P_b, g.shape[1],
                                            ptr(ls), ptr(os_), ptr(g), ptr(d_ls), ptr(d_os), ptr(d_diag),
                                            stream_ptr(g.device))
    check(rc, "lvae_kernel_blocks_bwd_f64")
    return d_ls, d_os, d_diag


def potrf_batched(A, check_info=True):
    """Lower Cholesky factors of a batch [..., n, n] (torch.cholesky, elbo_functions.py:177,179,185)."""
    lib = require_cuda(A)
    out = _c(A).clone()
    n = out.shape[-1]
    batch = out.numel() // (n * n)
    info = torch.zeros(1, dtype=torch.int32, device=out.device)
    with torch.cuda.device(out.device):
        rc = lib.lvae_potrf_batched_f64(ptr(out), n, n * n, batch, ptr(info), stream_ptr(out.device))
    check(rc, "lvae_potrf_batched_f64")
    if check_info and int(info.item()) != 0:
        raise RuntimeError(f"cholesky: matrix {int(info.item()) - 1} of the batch is not positive-definite")
    return out


def potri_batched(Lc):
    """Explicit inverse from a Cholesky factor (cholesky_solve(I, L), elbo_functions.py:178,180,186)."""
    lib = require_cuda(Lc)
    Lc = _c(Lc)
    n = Lc.shape[-1]
    batch = Lc.numel() // (n * n)
    out = torch.empty_like(Lc)
    with torch.cuda.device(Lc.device):
        rc = lib.lvae_potri_batched_f64(ptr(Lc), ptr(out), n, n * n, batch, stream_ptr(Lc.device))
    check(rc, "lvae_potri_batched_f64")
    return out


class KldCall:
    """One prepared call of the KL-bound op: owns outputs/scratch and the C struct (reusable across steps of one shape)."""

    def __init__(self, structure, L, M, Q, P_b, N_b, T_max, sum_T2, device, natural_gradient=True, path=0):
        self.lib = require_cuda()
        self.structure, self.device = structure, torch.device(device)
        self.L, self.M, self.Q, self.P_b, self.N_b = L, M, Q, P_b, N_b
        dev = self.device
        e = lambda *s: torch.empty(*s, dtype=F64, device=dev)
        self.kld_per_latent = e(L)
        self.grad_m, self.grad_H = e(L, M), e(L, M, M)
        self.d_mu, self.d_log_v = e(N_b, L), e(N_b, L)
        self.d_hyper = e(structure.n_ls + structure.n_comp + 1, L)      # d/d [lengthscales | outputscales | noise], one buffer
        self.d_lengthscale = self.d_hyper[:structure.n_ls]
        self.d_outputscale = self.d_hyper[structure.n_ls:structure.n_ls + structure.n_comp]
        self.d_noise = self.d_hyper[structure.n_ls + structure.n_comp]
        self.info = torch.zeros(4, dtype=torch.int32, device=dev)
        self.ks, self._keep = make_spec(structure)
        p = KldProblemT()
        p.L, p.M, p.Q, p.P_b, p.N_b, p.T_max, p.sum_T2 = L, M, Q, P_b, N_b, T_max, int(sum_T2)
        p.natural_gradient, p.path = int(bool(natural_gradient)), int(path)
        p.ks = self.ks
        self.stats_stride = int(self.lib.lvae_kld_stats_stride(M, structure.n_ls, structure.n_comp))
        self.stats = e(L, self.stats_stride)             # fully written by the reduce kernel of the subject pass
        self.workspace = e(int(self.lib.lvae_kld_workspace_doubles(C.byref(p))))
        for name in ("kld_per_latent", "grad_m", "grad_H", "d_mu", "d_log_v", "d_lengthscale", "d_outputscale",
                     "d_noise", "stats", "workspace", "info"):
            setattr(p, name, getattr(self, name).data_ptr())
        self.p = p
        self._held = ()
        off = int(self.lib.lvae_kld_hinv_offset(C.byref(p)))
        self.Hinv = self.workspace[off:off + L * M * M].view(L, M, M)      # H^-1 left by the head kernel

    def bind(self, x, offsets_dev, mu, log_v, z, m, H, lengthscale, outputscale, noise, scale, const_term, eps):
        held = [_c(t) for t in (x, mu, log_v, z, m, H, lengthscale, outputscale, noise)]
        p = self.p
        (p.x, p.mu, p.log_v, p.z, p.m, p.H, p.lengthscale, p.outputscale, p.noise) = [t.data_ptr() for t in held]
        p.offsets = offsets_dev.data_ptr()
        p.scale, p.const_term, p.eps = float(scale), float(const_term), float(eps)
        self._held = (held, offsets_dev)
        return self

    def set_stats(self, t):
        """Point the statistics row at `t` (e.g. a symmetric-memory region) for the next call."""
        self.p.stats = t.data_ptr()

    def _run(self, fn, what):
        with torch.cuda.device(self.device):
            check(fn(C.byref(self.p), stream_ptr(self.device)), what)

    def head(self):
        self._run(self.lib.lvae_kld_head_f64, "lvae_kld_head_f64")

    def subjects(self):
        self._run(self.lib.lvae_kld_subjects_f64, "lvae_kld_subjects_f64")

    def tail(self):
        self._run(self.lib.lvae_kld_tail_f64, "lvae_kld_tail_f64")

    def run(self):
        self._run(self.lib.lvae_kld_minibatch_f64, "lvae_kld_minibatch_f64")

    def post_info(self):
        """Deferred check: copy the flags to pinned host memory on the current stream without blocking."""
        if getattr(self, "_info_host", None) is None:
            self._info_host = torch.empty(4, dtype=torch.int32).pin_memory()
        self._info_host.copy_(self.info, non_blocking=True)
        self._info_event = torch.cuda.Event()
        self._info_event.record(torch.cuda.current_stream(self.device))

    def raise_on_info(self):
        ev = getattr(self, "_info_event", None)
        if ev is not None:                 # posted earlier: wait for THAT copy only, not for the work enqueued since
            ev.synchronize()
            info = self._info_host.tolist()
            self._info_event = None
        else:
            info = self.info.tolist()      # one device->host sync, like torch.cholesky's own check
        names = ("Kzz + eps*I", "H", "a per-subject block K1 + noise*I", "the natural-gradient update")
        for v, n in zip(info, names):
            if v:
                self.info.zero_()
                raise RuntimeError(f"cholesky: {n} is not positive-definite (flat index {v - 1})")


class SplitKldCall:
    """Ragged minibatches with M <= 62 whose longest subject has more than 24 rows: the subjects with at most 24 rows go
    through the 24-row-group instance of the fused kernels (one-warp prep tasks, two CTAs per SM), the longer ones through the
    40-row-group instance (four-warp prep tasks); both passes write their own statistics row, the rows are added (every
    batch-dependent statistic is a sum over subjects) and ONE tail runs.  Same interface as KldCall; d_mu / d_log_v come back
    in the caller's row order."""

    def __init__(self, structure, L, M, Q, counts, device, natural_gradient=True, path=0):
        counts = np.asarray(counts, dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(counts)])
        self.device, self.L, self.M, self.N_b = torch.device(device), L, M, int(off[-1])
        self.parts = []
        for sel in (counts <= 24, counts > 24):
            idx = np.nonzero(sel)[0]
            c = counts[idx]
            start = np.cumsum(c) - c                                      # first local row of every selected subject
            rows = np.repeat(off[idx] - start, c) + np.arange(int(c.sum()))   # global row of every local row
            call = KldCall(structure, L, M, Q, len(idx), int(c.sum()), int(c.max()), int((c * c).sum()), device,
                           natural_gradient, path)
            offs = torch.from_numpy(np.concatenate([[0], np.cumsum(c)]).astype(np.int32)).to(self.device)
            self.parts.append((call, torch.from_numpy(rows).to(self.device), offs))
        a = self.parts[0][0]
        self.kld_per_latent, self.grad_m, self.grad_H = a.kld_per_latent, a.grad_m, a.grad_H
        self.d_lengthscale, self.d_outputscale, self.d_noise, self.d_hyper = a.d_lengthscale, a.d_outputscale, a.d_noise, a.d_hyper
        self.stats, self.Hinv, self.info = a.stats, a.Hinv, a.info
        self.d_mu = torch.empty(self.N_b, L, dtype=F64, device=self.device)
        self.d_log_v = torch.empty(self.N_b, L, dtype=F64, device=self.device)
        self._target = self.stats

    def bind(self, x, offsets_dev, mu, log_v, z, m, H, lengthscale, outputscale, noise, scale, const_term, eps):
        for call, rows, offs in self.parts:
            call.bind(x[rows], offs, mu[rows], log_v[rows], z, m, H, lengthscale, outputscale, noise, scale, const_term, eps)
        return self

    def set_stats(self, t):
        self._target = t

    def head(self):
        for call, _, _ in self.parts:
            call.head()

    def subjects(self):
        (a, _, _), (b, _, _) = self.parts
        a.subjects()
        b.subjects()
        torch.add(a.stats, b.stats, out=self._target.view_as(a.stats) if self._target.numel() == a.stats.numel()
                  else self._target[:a.stats.numel()].view_as(a.stats))

    def tail(self):
        (a, ra, _), (b, rb, _) = self.parts
        a.tail()
        self.d_mu[ra], self.d_mu[rb] = a.d_mu, b.d_mu
        self.d_log_v[ra], self.d_log_v[rb] = a.d_log_v, b.d_log_v

    def run(self):
        self.head()
        self.subjects()
        self.tail()

    def post_info(self):
        for call, _, _ in self.parts:
            call.post_info()

    def raise_on_info(self):
        for call, _, _ in self.parts:
            call.raise_on_info()


class LatentTailKldCall:
    """Subjects sharded across ranks for the subject pass, LATENT dimensions sharded for everything that is per latent
    (SURVEY 8e): rank r runs head, tail and the natural-gradient update for its L / world latents only.

        head()      own latents: Kzz^-1, H^-1, a, G, W ; all-gather of (W, a) into the all-latent problem's workspace
        subjects()  all latents, this rank's subjects ; reduce-scatter of the statistics rows by latent
        tail()      own latents ; all-gather of kld per latent and of the hyper-parameter gradients (small)
        ng_step()   own latents ; all-gather of the new (m, H)

    With the replicated tail every rank repeats the O(L M^3) per-latent work, which caps strong scaling once M >= 128
    (cfg3: 2.4 of 23 ms per step do not shrink with the number of GPUs).  Same interface as KldCall; grad_m / grad_H are
    full-size tensors of which only this rank's latents are filled (the others are zero) — ng_step() consumes them."""

    def __init__(self, full, group):
        import torch.distributed as dist
        self.full, self.group = full, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L, M = full.L, full.M
        if L % self.world:
            raise RuntimeError(f"lvae_b200: the latent-sharded tail needs L ({L}) divisible by the number of ranks ({self.world})")
        self.L, self.M, self.Q, self.device, self.structure = L, M, full.Q, full.device, full.structure
        self.Lo = L // self.world
        self.l0, self.l1 = self.rank * self.Lo, (self.rank + 1) * self.Lo
        ng = bool(full.p.natural_gradient)
        self.own = KldCall(full.structure, self.Lo, M, full.Q, 0, 0, int(full.p.T_max), 0, full.device, ng, int(full.p.path))
        lib = full.lib
        offs = (C.c_int64 * 4)()
        check(lib.lvae_kld_head_offsets(C.byref(full.p), offs), "lvae_kld_head_offsets")
        self.wstr = int(offs[1])
        self.W_full = full.workspace[int(offs[0]):int(offs[0]) + L * self.wstr].view(L, self.wstr)
        self.a_full = full.workspace[int(offs[2]):int(offs[2]) + L * M].view(L, M)
        check(lib.lvae_kld_head_offsets(C.byref(self.own.p), offs), "lvae_kld_head_offsets")
        if int(offs[1]) != self.wstr:
            raise RuntimeError("lvae_b200: the per-latent and the all-latent problem disagree on the kernel path")
        self.W_own = self.own.workspace[int(offs[0]):int(offs[0]) + self.Lo * self.wstr].view(self.Lo, self.wstr)
        self.a_own = self.own.workspace[int(offs[2]):int(offs[2]) + self.Lo * M].view(self.Lo, M)
        e = lambda *sh: torch.empty(*sh, dtype=F64, device=full.device)
        self.pack_own, self.pack_all = e(self.Lo, self.wstr + M), e(L, self.wstr + M)
        nh = full.d_hyper.shape[0]
        self.small_own, self.small_all = e(self.Lo, nh + 1), e(L, nh + 1)
        self.kld_per_latent, self.d_hyper = e(L), e(nh, L)
        self.d_lengthscale = self.d_hyper[:self.structure.n_ls]
        self.d_outputscale = self.d_hyper[self.structure.n_ls:self.structure.n_ls + self.structure.n_comp]
        self.d_noise = self.d_hyper[self.structure.n_ls + self.structure.n_comp]
        self.grad_m = torch.zeros(L, M, dtype=F64, device=full.device)
        self.grad_H = torch.zeros(L, M, M, dtype=F64, device=full.device)
        self.own.p.grad_m = self.grad_m[self.l0:self.l1].data_ptr()        # the own tail writes straight into the slices
        self.own.p.grad_H = self.grad_H[self.l0:self.l1].data_ptr()
        self.d_mu, self.d_log_v, self.stats, self.info, self.Hinv = full.d_mu, full.d_log_v, full.stats, full.info, self.own.Hinv
        self._empty_off = torch.zeros(1, dtype=torch.int32, device=full.device)
        self.ng_ws = e(int(lib.lvae_ng_workspace_doubles(self.Lo, M)))
        self.ng_info = torch.zeros(4, dtype=torch.int32, device=full.device)

    def bind(self, x, offsets_dev, mu, log_v, z, m, H, lengthscale, outputscale, noise, scale, const_term, eps):
        self.full.bind(x, offsets_dev, mu, log_v, z, m, H, lengthscale, outputscale, noise, scale, const_term, eps)
        l0, l1 = self.l0, self.l1
        self.own.bind(x[:0], self._empty_off, mu[:0, l0:l1], log_v[:0, l0:l1], z[l0:l1], m[l0:l1], H[l0:l1],
                      lengthscale[:, l0:l1], outputscale[:, l0:l1], noise[l0:l1], scale, const_term * self.Lo / self.L, eps)
        return self

    def set_stats(self, t):
        self.full.set_stats(t)

    def head(self):
        import torch.distributed as dist
        self.own.head()
        self.pack_own[:, :self.wstr].copy_(self.W_own)
        self.pack_own[:, self.wstr:].copy_(self.a_own)
        dist.all_gather_into_tensor(self.pack_all, self.pack_own, group=self.group)       # W, a of every latent
        self.W_full.copy_(self.pack_all[:, :self.wstr])
        self.a_full.copy_(self.pack_all[:, self.wstr:])

    def subjects(self):
        import torch.distributed as dist
        self.full.subjects()
        dist.reduce_scatter_tensor(self.own.stats, self.full.stats, group=self.group)     # statistics rows, by latent

    def tail(self):
        import torch.distributed as dist
        self.own.tail()
        self.small_own[:, 0].copy_(self.own.kld_per_latent)
        self.small_own[:, 1:].copy_(self.own.d_hyper.t())
        dist.all_gather_into_tensor(self.small_all, self.small_own, group=self.group)
        self.kld_per_latent.copy_(self.small_all[:, 0])
        self.d_hyper.copy_(self.small_all[:, 1:].t())

    def run(self):
        self.head()
        self.subjects()
        self.tail()

    def ng_step(self, m, H, lr, gather=True):
        """training.py:129-135 on this rank's latents, in place in m [L,M(,1)] and H [L,M,M]; then the new (m, H) of every
        latent are all-gathered (gather=False leaves the other ranks' slices stale)."""
        import torch.distributed as dist
        L, M, l0, l1 = self.L, self.M, self.l0, self.l1
        mv = m.view(L, M)
        with torch.cuda.device(self.device):
            check(self.full.lib.lvae_ng_step_f64(ptr(mv[l0:l1]), ptr(H[l0:l1]), ptr(self.grad_m[l0:l1]), ptr(self.grad_H[l0:l1]),
                                                 ptr(self.own.Hinv), float(lr), self.Lo, M, ptr(self.ng_ws), ptr(self.ng_info),
                                                 stream_ptr(self.device)), "lvae_ng_step_f64")
        if gather:
            dist.all_gather_into_tensor(mv, mv[l0:l1].clone(), group=self.group)
            dist.all_gather_into_tensor(H, H[l0:l1].clone(), group=self.group)

    def post_info(self):
        self.full.post_info()
        self.own.post_info()

    def raise_on_info(self):
        self.full.raise_on_info()
        self.own.raise_on_info()


def make_kld_call(structure, L, M, Q, counts, device, natural_gradient=True, path=0, split=True):
    """KldCall for a minibatch whose subjects have `counts` rows (host array), or SplitKldCall when that is faster
    (split=False: always one call)."""
    counts = np.asarray(counts, dtype=np.int64)
    if split and path != 1 and M <= 62 and counts.size and counts.max() > 24:
        short = counts <= 24
        if short.any() and counts[short].sum() >= 0.1 * counts.sum():
            return SplitKldCall(structure, L, M, Q, counts, device, natural_gradient, path)
    return KldCall(structure, L, M, Q, counts.size, int(counts.sum()), int(counts.max()) if counts.size else 0,
                   int((counts * counts).sum()), device, natural_gradient, path)


def ng_step(m, H, grad_m, grad_H, lr, Hinv=None):
    """Natural-gradient update of (m [L,M,1], H [L,M,M]) — training.py:129-135.  Returns new detached tensors.
    Hinv: H^-1 already computed by the bound's head kernel for this H (saves one Cholesky + inverse)."""
    lib = require_cuda(m, H, grad_m, grad_H)
    L, M = H.shape[0], H.shape[-1]
    m2, H2 = _c(m).clone(), _c(H).clone()
    gm, gH = _c(grad_m), _c(grad_H)
    ws = torch.empty(int(lib.lvae_ng_workspace_doubles(L, M)), dtype=F64, device=H.device)
    info = torch.zeros(4, dtype=torch.int32, device=H.device)
    hi = _c(Hinv) if Hinv is not None else None
    with torch.cuda.device(H.device):
        rc = lib.lvae_ng_step_f64(ptr(m2), ptr(H2), ptr(gm), ptr(gH), ptr(hi), float(lr), L, M, ptr(ws), ptr(info),
                                  stream_ptr(H.device))
    check(rc, "lvae_ng_step_f64")
    return m2.view_as(m), H2, info


def launch_count():
    return int(_lib.load().lvae_launch_count())


def gemm_batched(A, B, trans_a=False, trans_b=False, alpha=1.0, beta=0.0, C=None, flags=0):
    """C[b] = alpha * op(A[b]) op(B[b]) + beta * C[b] on the FP64 tensor pipe (lvae_gemm_batched_f64).  A, B: [batch, r, c]
    contiguous.  flags: 1 = lower triangle only, 3 = lower computed and mirrored (symmetric results)."""
    lib = require_cuda(A, B)
    A, B = _c(A), _c(B)
    batch = A.shape[0]
    m, k = (A.shape[2], A.shape[1]) if trans_a else (A.shape[1], A.shape[2])
    n = B.shape[1] if trans_b else B.shape[2]
    if C is None:               # every entry is written unless only the lower triangle is asked for (no memset of 300 MB outputs)
        full = beta == 0.0 and (int(flags) & 3) != 1 and min(m, n) > 0
        C = (torch.empty if full else torch.zeros)(batch, m, n, dtype=F64, device=A.device)
    with torch.cuda.device(A.device):
        rc = lib.lvae_gemm_batched_f64(int(trans_a), int(trans_b), m, n, k, float(alpha), ptr(A), A.shape[2],
                                       A.shape[1] * A.shape[2], ptr(B), B.shape[2], B.shape[1] * B.shape[2], float(beta),
                                       ptr(C), n, m * n, batch, int(flags), stream_ptr(A.device))
    check(rc, "lvae_gemm_batched_f64")
    return C
