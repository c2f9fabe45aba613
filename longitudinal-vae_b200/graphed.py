"""The whole GP side of one Hensman step as ONE CUDA-graph replay, for fixed-shape minibatches.

At the reference's default of 20 subjects per minibatch (parse_model_args.py:94) the GPU work of a step is ~0.2 ms while
the Python around it — flattening the kernel modules, the packed hyper-parameter transform, ~10 ctypes calls, autograd
bookkeeping — costs 0.6-0.7 ms.  `GraphedHensmanStep` captures, once, everything that does not depend on Python state:

    hyper-parameter transform (softplus + bounds)  ->  lvae_kld_minibatch_f64 (bound + every adjoint)
    ->  chain rule back to the raw hyper-parameters  ->  lvae_ng_step_f64 on (m, H) in place (training.py:129-135)

and replays it per step on static buffers.  What stays outside is what the reference's loop does around the bound
(training.py:103-127): the VAE forward, `net_loss.backward()` and `optimiser.step()` — the returned `kld` is an autograd
node w.r.t. `mu`, `log_var` and the raw kernel / noise parameters, whose gradients were already computed by the replay and
are only scaled in `backward`.

Restrictions: natural-gradient mode, a fixed number of rows per subject, a fixed number of subjects per call (run the last,
shorter minibatch of an epoch through `elbo_functions.minibatch_KLD_upper_bound`), one process (no statistics exchange inside
the graph), and at most one step in flight: call `backward` on the returned value before the next call.
"""
import torch

from . import _lib, ops
from . import elbo_functions as EF
from .spec import build_structure, flatten

F64 = torch.float64


class _GraphedBound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, log_v, step, *params):
        ctx.step, ctx.dtypes = step, (mu.dtype, log_v.dtype)
        ctx.serial = step.serial
        return step.kld.clone()

    @staticmethod
    def backward(ctx, g):
        s = ctx.step
        if ctx.serial != s.serial:
            raise RuntimeError("lvae_b200: GraphedHensmanStep was called again before backward() of the previous step")
        outs = torch._foreach_mul([s.call.d_mu, s.call.d_log_v] + s.raw_grads, g)
        d_mu, d_lv = outs[0].to(ctx.dtypes[0]), outs[1].to(ctx.dtypes[1])
        return (d_mu, d_lv, None) + tuple(o.to(p.dtype) for o, p in zip(outs[2:], s.params))


class GraphedHensmanStep:
    """step = GraphedHensmanStep(covar_module0, covar_module1, likelihoods, latent_dim, m, H, zt_list, P, subjects_per_batch,
    T, eps, natural_gradient_lr);  kld_loss = step(train_x, mu, log_var)  replaces lines 108-115 and 129-135 of training.py.
    `m` [L,M,1] and `H` [L,M,M] (FP64, CUDA, contiguous) are updated IN PLACE by every call."""

    def __init__(self, covar_module0, covar_module1, likelihood, latent_dim, m, H, z, P_tot, P_batch, T, eps=1e-6,
                 natural_gradient_lr=0.01, Q=None):
        lib = _lib.require_cuda(m, H, z)
        if EF._GROUP is not None:
            raise RuntimeError("lvae_b200: GraphedHensmanStep is single-process (no statistics exchange inside the graph)")
        if m.dtype is not F64 or H.dtype is not F64 or not m.is_contiguous() or not H.is_contiguous():
            raise RuntimeError("lvae_b200: GraphedHensmanStep updates m and H in place: they must be contiguous FP64")
        self.lib, self.dev = lib, H.device
        self.cm0, self.cm1, self.lik = covar_module0, covar_module1, likelihood
        L, M = latent_dim, H.shape[-1]
        self.L, self.M, self.T, self.P_b = L, M, int(T), int(P_batch)
        self.m, self.H, self.lr = m, H, float(natural_gradient_lr)
        z = z.detach().to(F64)
        self.z = (z.unsqueeze(0).expand(L, -1, -1) if z.dim() == 2 else z).contiguous()
        Q = self.z.shape[-1] if Q is None else Q
        N_b = self.P_b * self.T
        dev = self.dev
        self.x = torch.zeros(N_b, Q, dtype=F64, device=dev)
        self.mu = torch.zeros(N_b, L, dtype=F64, device=dev)
        self.lv = torch.zeros(N_b, L, dtype=F64, device=dev)
        self.offsets = torch.arange(0, N_b + 1, self.T, dtype=torch.int32, device=dev)
        self.scale, self.const, self.eps = P_tot / P_batch, latent_dim * P_tot * T / 2, float(eps)
        st = self._hyper()[0]
        self.call = ops.KldCall(st, L, M, Q, self.P_b, N_b, self.T, self.P_b * self.T * self.T, dev, natural_gradient=True,
                                path=EF._PATH)
        self.params = [p for mod in (covar_module0, covar_module1, likelihood) for p in mod.parameters() if p.requires_grad]
        self.params = list({id(p): p for p in self.params}.values())
        self.raw_grads = [torch.zeros_like(p, dtype=F64) for p in self.params]
        self.ng_ws = torch.empty(int(lib.lvae_ng_workspace_doubles(L, M)), dtype=F64, device=dev)
        self.ng_info = torch.zeros(4, dtype=torch.int32, device=dev)
        self.kld = torch.zeros((), dtype=F64, device=dev)
        self.serial = 0
        self._pending = False
        self._capture()

    # -- graph body ---------------------------------------------------------------------------------------------------
    def _hyper(self):
        L = self.L
        st, ls, os_, nz = build_structure(flatten(self.cm0), flatten(self.cm1), L, device=self.dev,
                                          extra=[EF._noise_entry(self.lik)])
        same = ls._base is not None and ls._base is os_._base and ls._base is nz._base and \
            ls._base.shape[0] == st.n_ls + st.n_comp + 1
        hyper = ls._base if same else torch.cat([ls, os_, nz.reshape(1, L).to(F64)])
        return st, hyper

    def _body(self):
        L, M, c = self.L, self.M, self.call
        with torch.enable_grad():
            st, hyper = self._hyper()
        h = hyper.detach()
        c.bind(self.x, self.offsets, self.mu, self.lv, self.z, self.m.view(L, M), self.H, h[:st.n_ls],
               h[st.n_ls:st.n_ls + st.n_comp], h[st.n_ls + st.n_comp], self.scale, self.const, self.eps)
        c.run()
        if self.params and hyper.requires_grad:
            grads = torch.autograd.grad([hyper], self.params, [c.d_hyper], allow_unused=True)
            for buf, g in zip(self.raw_grads, grads):
                if g is not None:
                    buf.copy_(g)
        with torch.cuda.device(self.dev):
            rc = self.lib.lvae_ng_step_f64(_lib.ptr(self.m), _lib.ptr(self.H), _lib.ptr(c.grad_m), _lib.ptr(c.grad_H),
                                           _lib.ptr(c.Hinv), self.lr, L, M, _lib.ptr(self.ng_ws), _lib.ptr(self.ng_info),
                                           _lib.stream_ptr(self.dev))
        _lib.check(rc, "lvae_ng_step_f64")
        self.kld.copy_(c.kld_per_latent.sum())

    def _capture(self):
        m0, H0 = self.m.clone(), self.H.clone()
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):                       # warm-up outside the capture (lazy CUDA attributes, allocator)
            for _ in range(2):
                self._body()
                self.m.copy_(m0)
                self.H.copy_(H0)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.call.info.zero_()
        self.ng_info.zero_()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        self.m.copy_(m0)                                   # capturing executes nothing, but be explicit about the state
        self.H.copy_(H0)

    # -- per step -----------------------------------------------------------------------------------------------------
    def check_errors(self):
        """Raise if a Cholesky factorisation of an earlier step failed (flags are read without stalling the stream)."""
        if self._pending:
            self._pending = False
            self.call.raise_on_info()
            if int(self._ng_host[3]) != 0:
                raise RuntimeError(f"cholesky: natural-gradient update of latent {int(self._ng_host[3]) - 1} is not "
                                   "positive-definite")

    def __call__(self, train_x, mu, log_var):
        if train_x.shape[0] != self.x.shape[0]:
            raise RuntimeError(f"lvae_b200: GraphedHensmanStep was built for {self.x.shape[0]} rows, got {train_x.shape[0]}")
        self.check_errors()                                 # flags of the previous step: copied long ago, no stall
        self.x.copy_(train_x, non_blocking=True)
        self.mu.copy_(mu.detach(), non_blocking=True)
        self.lv.copy_(log_var.detach(), non_blocking=True)
        self.graph.replay()
        self.serial += 1
        if not hasattr(self, "_ng_host"):
            self._ng_host = torch.empty(4, dtype=torch.int32).pin_memory()
        self._ng_host.copy_(self.ng_info, non_blocking=True)
        self.call.post_info()                               # records the event both flag copies are waited on
        self._pending = True
        return _GraphedBound.apply(mu, log_var, self, *self.params)
