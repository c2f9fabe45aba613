#!/usr/bin/env python
"""bench.py — GP-prior ELBO-step throughput (subjects/sec) of the L-VAE hot path on B200.

    python bench.py --gpus N --steps K --warmup W              (N > 1: launched by torch.distributed.run, one rank/GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W     (the CPU reference arm, rank 0 only)

A "step" is one Hensman minibatch of the path: kernel construction from covariates, batched Cholesky/inverse of the
per-subject blocks and of K_mm/H, the minibatch KL upper bound, ALL its gradients (mu, log_v, kernel hyper-parameters,
noise) and the natural-gradient update of (m, H)  (elbo_functions.py:144-216 + backward + training.py:129-135).
Headline workload = BASELINE.json configs[1] (cfg2): Health-MNIST-shaped synthetic data, 1000 subjects x 20 time points
per GPU, L=32, M=60, full additive kernel (cat(id) + SE(age) + id x age + gender x age + disease x disease_time), all
subjects of a rank in one minibatch per step.  N>1: weak scaling — every rank holds its own 1000 subjects of a N*1000
-subject minibatch and the SVGP sufficient statistics are summed over ranks between the subject pass and the tail.

`value` is device-resident throughput (inputs in HBM, CUDA events, max over ranks); `e2e` is the same step through the
public Python API (lvae_b200.elbo_functions.minibatch_KLD_upper_bound + backward + natural_gradient_step) with pinned
HOST inputs copied in and the loss and encoder gradients copied out inside the timed region.  `parity` compares the
outputs of the CUDA path ON THE TIMED PROBLEM with the oracle port of the reference run with stock torch ops on the same
GPU (and, for N > 1, the sharded result with a one-GPU evaluation of the gathered minibatch); a failed check exits 1.
`other_configs` carries the same measurements for BASELINE configs[2..4] (cfg3, cfg4, cfg5) and a strong-scaling point.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

if "--impl" in sys.argv and "reference" in sys.argv[sys.argv.index("--impl") + 1:sys.argv.index("--impl") + 2] or \
        "--impl=reference" in sys.argv:
    # The reference picks `torch.device("cuda" if torch.cuda.is_available() else "cpu")` inside every function
    # (elbo_functions.py:165, 240): the CPU arm must not see the GPU of the box it runs on.
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gp_prior_elbo_step_subjects_per_sec"
UNIT = "subjects/s"
LR = 1e-3
EPS = 1e-6
# Parity tolerances, max-norm relative per output tensor (north_star: 1e-6 FP64).  The K0 hyper-parameter gradients are
# differences of terms ~1/eps larger than the result when cond(Kzz + eps I) ~ 1e8: the hyper-
# gradient VECTOR is compared per latent-stacked vector at 1e-6 of its max-norm like the other tensors, and its
# worst single entry is reported next to it (not gated).
TOL = 1e-6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cfg", default="cfg2")
    ap.add_argument("--spb", type=int, default=None, help="subjects per minibatch per GPU (default: per config)")
    ap.add_argument("--L", type=int, default=None)
    ap.add_argument("--M", type=int, default=None)
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 generic kernels, 2 fused DMMA kernel")
    ap.add_argument("--cpu-subjects", type=int, default=None,
                    help="subjects in the CPU arm's per-step sample (default: the whole minibatch of one rank for fixed T "
                         "and M <= 64, else 100)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["nccl", "p2p"],
                    help="N > 1: statistics exchange by NCCL all-reduce or by the peer-memory kernel (symmetric memory)")
    ap.add_argument("--tail", default="auto", choices=["auto", "replicated", "latents"],
                    help="N > 1: head / tail / NG step replicated on every rank, or sharded by latent (reduce-scatter of the "
                         "statistics by latent + all-gather of W, a); auto = latents when M > 64 or for strong scaling")
    ap.add_argument("--no-split", action="store_true", help="ragged minibatches: one call with 40-row groups for all subjects "
                                                              "instead of splitting them by length (measurement)")
    ap.add_argument("--no-latency-point", action="store_true", help="skip the spb=20 point (the reference's default batch)")
    ap.add_argument("--no-others", action="store_true", help="skip cfg3 / cfg4 / cfg5 and the strong-scaling point")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--others-steps", type=int, default=5)
    ap.add_argument("--profile-e2e", default=None, help="diagnostic: torch.profiler tables of three API steps -> PATH.<cfg>.txt")
    return ap.parse_args()


DEFAULT_SPB = {"cfg1": 100, "cfg2": 1000, "cfg3": 1000, "cfg4": 2500, "cfg5": 2000}


# ---------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------
def make_problem(cfg, spb, rank, world, L=None, M=None):
    from lvae_b200 import synth
    b = synth.make_batch(cfg, P=spb, L=L, M=M, seed=1234 + int(cfg[-1]) + 1000 * rank, first_subject=rank * spb)
    if world > 1:   # inducing points and (m, H) are replicated: take rank 0's
        b0 = synth.make_batch(cfg, P=spb, L=L, M=M, seed=1234 + int(cfg[-1]))
        b.z, b.m, b.H = b0.z, b0.m, b0.H
    return b


def n_components(lists):
    return (len(lists["cat_kernel"]) + len(lists["sqexp_kernel"]) + len(lists["bin_kernel"]) + len(lists["cat_int_kernel"]) +
            len(lists["bin_int_kernel"]))


def workload_config(cfg, spb, world, b, strong=False):
    """The `config` object — identical in both arms (ours and --impl reference) for the same flags."""
    Tdesc = f"T={b.T}" if not isinstance(b.T, tuple) else f"T in {b.T[0]}..{b.T[1]} (ragged)"
    glob = spb if strong else spb * world
    return {"workload": f"{cfg}: Health-MNIST-shaped synthetic, {Tdesc}, L={b.L}, M={b.M}, additive kernel with "
                        f"{n_components(b.lists)} components, one minibatch of {glob} subjects per step over {world} GPU(s) "
                        f"(bound + all gradients + natural-gradient update), perturbed hyper-parameters",
            "global_batch_subjects": glob, "L": b.L, "M": b.M, "n_gpus": world}


def hyper_values(b):
    """Perturbed hyper-parameters of SURVEY 8(d) (so that latents differ): (lengthscale [n_ls,L], outputscale [n_comp,L],
    noise [L]) in K0-then-K1 component order."""
    from lvae_b200 import synth
    lists = b.lists
    id_cov = 2
    n_ls = len(lists["sqexp_kernel"]) + len(lists["cat_int_kernel"]) + len(lists["bin_int_kernel"])
    with torch.device("cpu"):                     # seeded CPU generator, also when called under a CUDA device context
        ls, os_, noise = synth.perturbed_hypers(n_ls, n_components(lists), b.L, seed=1234, noise_trainable=True)
        mode = os.environ.get("LVAE_BENCH_HYPERS", "perturbed")      # diagnostics: "default" | "unit_noise"
        if mode == "default":
            ls, os_, noise = torch.full_like(ls, 2.5), torch.full_like(os_, float(np.log(2.0))), torch.ones_like(noise)
        elif mode == "unit_noise":
            noise = torch.ones_like(noise)
    return ls, os_, noise, id_cov


def algorithmic_flops(T, L, M, C0, C1, P_b):
    """FP64 flops the algorithm needs (FMA = 2), un-padded — DESIGN.md 'Kernels and rooflines'.
    subject kernel per (subject, latent): V = Bi Kxz (2T^2M), S += Kxz^T V (2TM^2), Y = V W (2TM^2), Q = Y V^T (2T^2M),
    kernel entries and their adjoint contractions (8 flop per component entry incl. exp, SURVEY 8d convention, + 4 per
    entry for the two hyper-gradient dot products), r, u, ng1, da."""
    T = np.asarray(T, dtype=np.float64)
    subj = (4 * T * M * M + 4 * T * T * M + 12 * C0 * T * M + 12 * C1 * T * T + 8 * T * M + 4 * T * T).sum() * L
    prep = ((7.0 / 3) * T ** 3 + 4 * T ** 3 + 12 * (C0 + C1) * T * T).sum() * L
    fixed = L * ((38.0 / 3) + 10 + (7.0 / 3) * 2) * M ** 3      # head (chol x2, inverse x2, G) + tail GEMMs + NG step
    return dict(subjects=float(subj), prep=float(prep), fixed=float(fixed), step=float(subj + prep + fixed))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop, self.thr = index, [], threading.Event(), None

    def _loop(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        self.thr = threading.Thread(target=self._loop, daemon=True)
        self.thr.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thr.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dgemm_peak_tflops(device):
    """cuBLAS FP64 GEMM burst peak on this GPU (MEASURED_PEAKS.json has no FP64 entry): 4096^3, best of 5."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize(device)
        best = min(best, e0.elapsed_time(e1))
    return 2 * n ** 3 / best * 1e-9


# ---------------------------------------------------------------------------------------------------------------
# The oracle port of the reference (oracle/lvae_oracle.py): the CHECKER of `parity`, the "existing GPU path" when run with
# stock torch CUDA ops, and the CPU arm's fallback when baseline/_ref is absent.
# ---------------------------------------------------------------------------------------------------------------
def _oracle():
    p = os.path.join(ROOT, "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)
    import lvae_oracle as orc
    return orc


def oracle_components(b, device, requires_grad=True, latents=None):
    """Oracle kernel components with this problem's hyper-parameters on `device`; returns (k0, k1, noise, params) with
    params ordered [lengthscale rows | outputscale rows | noise] like the library's d_hyper.  latents: index list (the
    latent dimensions of the bound are independent, so a subset is a smaller instance of the same problem)."""
    orc = _oracle()
    ls, os_, noise, id_cov = hyper_values(b)
    if latents is not None:
        ls, os_, noise = ls[:, latents], os_[:, latents], noise[latents]
    k0, k1 = orc.parse_kernel_lists(len(latents) if latents is not None else b.L, **b.lists, id_covariate=id_cov)
    ls_p, os_p = [], []
    i_ls = 0
    for i_c, comp in enumerate(k0 + k1):
        comp.outputscale = os_[i_c].to(device).clone().requires_grad_(requires_grad)
        os_p.append(comp.outputscale)
        for k in sorted(comp.lengthscales):
            comp.lengthscales[k] = ls[i_ls].to(device).clone().requires_grad_(requires_grad)
            ls_p.append(comp.lengthscales[k])
            i_ls += 1
    nz = noise.to(device).clone().requires_grad_(requires_grad)
    return k0, k1, nz, ls_p + os_p + [nz]


def oracle_step_fn(b, n_subjects, device="cpu", latents=None, eps=None):
    """One full step of the oracle port on the first n_subjects of b, taken as the whole data set (P_tot = P_batch =
    n_subjects, so the minibatch scale is 1): fwd + backward + NG update.  Returns step() -> dict of outputs (kld, grad_m,
    grad_H, d_mu, d_log_v, d_hyper [n_hyp, L]).  latents: restrict to these latent dimensions; eps: jitter override."""
    orc = _oracle()
    L = b.L if latents is None else len(latents)
    rows = int(b.offsets[n_subjects])
    P_tot = P_glob = n_subjects
    N_tot = rows
    dev = lambda t: t.to(device)
    sel = (lambda t, d: t) if latents is None else (lambda t, d: t.index_select(d, torch.as_tensor(latents, device=t.device)))
    x, mu0, lv0, z = dev(b.x[:rows]), dev(sel(b.mu[:rows], 1)), dev(sel(b.log_v[:rows], 1)), dev(sel(b.z, 0))
    k0, k1, noise, params = oracle_components(b, device, latents=latents)
    ragged = isinstance(b.T, tuple)
    state = {"m": dev(sel(b.m, 0)).clone(), "H": dev(sel(b.H, 0)).clone()}
    EPS = globals()["EPS"] if eps is None else eps

    def step(update=True):
        mu = mu0.clone().requires_grad_(True)
        lv = lv0.clone().requires_grad_(True)
        if ragged:
            kld, gm, gH = orc.kld_iter(k0, k1, noise, L, state["m"], state["H"], x, mu, lv, z, P_tot, P_glob, N_tot,
                                       True, 2, EPS)
        else:
            kld, gm, gH = orc.kld_fixed_T(k0, k1, noise, L, state["m"], state["H"], x, mu, lv, z, P_tot, P_glob, int(b.T),
                                          True, EPS)
        kld.sum().backward()
        out = dict(kld=kld.detach().sum(), grad_m=gm.detach(), grad_H=gH.detach(), d_mu=mu.grad, d_log_v=lv.grad,
                   d_hyper=torch.stack([p_.grad for p_ in params]))
        if update:
            state["m"], state["H"] = orc.ng_step(state["m"], state["H"], gm.detach(), gH.detach(), LR)
        out["m_new"], out["H_new"] = (state["m"], state["H"]) if update else (None, None)
        for p_ in params:
            p_.grad = None
        return out
    return step


@contextlib.contextmanager
def reference_on_cpu():
    """The reference chooses its device with torch.cuda.is_available() at call time (elbo_functions.py:165, 240); inside
    this context that answer is False, so its unmodified code runs on the host cores of a box that has a GPU."""
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        yield
    finally:
        torch.cuda.is_available = real


def reference_step_fn(b, n_subjects):
    """The reference's OWN functions (byte-for-byte copies under baseline/_ref, oracle/install_ref.py) on the host CPU:
    kernel_gen.generate_kernel_batched + elbo_functions.minibatch_KLD_upper_bound[_iter] + backward + the natural-gradient
    update of training.py:129-135.  GPyTorch (absent third party) resolves to oracle/gpytorch_standin."""
    import importlib.util
    import warnings
    warnings.filterwarnings("ignore")
    orc = _oracle()
    sys.path.insert(0, os.path.join(ROOT, "oracle", "gpytorch_standin"))
    import gpytorch
    ref = {}
    for name in ("kernel_spec", "kernel_gen", "elbo_functions"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "baseline", "_ref", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        ref[name] = mod
    L = b.L
    ls, os_, noise, id_cov = hyper_values(b)
    with reference_on_cpu():                      # kernel_gen.py:219 moves the modules to "cuda" when it is available
        cm0, cm1 = ref["kernel_gen"].generate_kernel_batched(L, **b.lists, id_covariate=id_cov)
    cm0.double(), cm1.double()
    i_c = i_l = 0
    for mod in (cm0, cm1):
        for sk in mod.kernels:
            sk.outputscale = os_[i_c].clone()
            i_c += 1
            for rb in [mm for mm in sk.modules() if isinstance(mm, gpytorch.kernels.RBFKernel)]:
                rb.lengthscale = ls[i_l].clone().view(-1, 1, 1)
                i_l += 1
    lik = gpytorch.likelihoods.GaussianLikelihood(batch_shape=torch.Size([L]),
                                                  noise_constraint=gpytorch.constraints.GreaterThan(1e-8)).double()
    lik.noise = noise.view(L, 1)
    rows = int(b.offsets[n_subjects])
    P_tot = P_glob = n_subjects                 # the sample is the whole data set of the call (minibatch scale 1)
    N_tot = rows
    x, mu0, lv0 = b.x[:rows], b.mu[:rows], b.log_v[:rows]
    ragged = isinstance(b.T, tuple)
    state = {"m": b.m.clone(), "H": b.H.clone()}
    EFr = ref["elbo_functions"]

    def step():
        mu = mu0.clone().requires_grad_(True)
        lv = lv0.clone().requires_grad_(True)
        with reference_on_cpu():
            if ragged:
                kld, gm, gH = EFr.minibatch_KLD_upper_bound_iter(cm0, cm1, lik, L, state["m"], state["H"], x, mu, lv, b.z,
                                                                 P_tot, P_glob, N_tot, True, id_cov, EPS)
            else:
                kld, gm, gH = EFr.minibatch_KLD_upper_bound(cm0, cm1, lik, L, state["m"], state["H"], x, mu, lv, b.z, P_tot,
                                                            P_glob, int(b.T), True, EPS)
            kld.sum().backward()
        state["m"], state["H"] = orc.ng_step(state["m"], state["H"], gm.detach(), gH.detach(), LR)   # training.py:129-135
        for mod in (cm0, cm1, lik):
            mod.zero_grad(set_to_none=True)
        return dict(kld=kld.detach().sum())
    return step


def time_cpu(b, n_subjects, steps, warmup, prefer_ref=True):
    """CPU arm: the reference's own modules when baseline/_ref holds them, else the oracle port; all host threads."""
    torch.set_num_threads(os.cpu_count() or 1)
    kind = "port"
    step = None
    if prefer_ref and all(os.path.exists(os.path.join(ROOT, "baseline", "_ref", f)) for f in
                          ("elbo_functions.py", "kernel_gen.py", "kernel_spec.py")):
        try:
            step = reference_step_fn(b, n_subjects)
            kind = "reference"
        except Exception as ex:                       # e.g. a torch API the reference needs is gone: say so, use the port
            print(f"bench: baseline/_ref could not be driven ({repr(ex)[:200]}); using the oracle port", file=sys.stderr)
    if step is None:
        step = oracle_step_fn(b, n_subjects)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n_subjects / dt, dt, kind


def default_cpu_subjects(args, b, spb):
    if args.cpu_subjects is not None:
        return min(args.cpu_subjects, spb)
    return spb if (not isinstance(b.T, tuple) and b.M <= 64 and spb <= 1000) else min(100, spb)


def time_gpu_reference(b, device, steps=3, warmup=2):
    """The oracle port of the reference with stock PyTorch CUDA ops (ATen / cuBLAS / cuSOLVER) on this GPU: the "existing
    GPU path" (SURVEY 8d).  All subjects of the rank's minibatch, fwd + backward + NG update."""
    with torch.device(device):
        step = oracle_step_fn(b, b.P, device=device)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / steps
    return b.P / (ms * 1e-3), ms


def rel_err(a, ref):
    a, ref = a.detach().double().reshape(-1), ref.detach().double().reshape(-1).to(a.device)
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-300))


# ---------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner) write to fd 1; the contract is ONE JSON line there.  Park fd 1 on stderr and keep
    the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ---------------------------------------------------------------------------------------------------------------
# one configuration on this rank's GPU
# ---------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def zero_grads(params):
    """What optimiser.zero_grad(set_to_none=True) does (training.py:105 of the reference): the parameter list is collected
    once, not by walking the module trees every step (Module.zero_grad: 0.26 ms per step for the three GP modules)."""
    for p in params:
        p.grad = None


def build_modules(b, device):
    """Drop-in kernel modules + likelihood of this package with the problem's hyper-parameters."""
    from lvae_b200.constraints import GreaterThan
    from lvae_b200.gp_kernels import RBFKernel
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200.likelihoods import GaussianLikelihood
    ls, os_, noise, id_cov = hyper_values(b)
    L = b.L
    cm0, cm1 = generate_kernel_batched(L, **b.lists, id_covariate=id_cov)
    cm0, cm1 = cm0.double().to(device), cm1.double().to(device)
    i_c = i_l = 0
    for mod in (cm0, cm1):
        for sk in mod.kernels:
            sk.outputscale = os_[i_c]
            i_c += 1
            for rb in [mm for mm in sk.modules() if isinstance(mm, RBFKernel)]:
                rb.lengthscale = ls[i_l].view(L, 1, 1)
                i_l += 1
    lik = GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=GreaterThan(1e-8)).double().to(device)
    lik.noise = noise.view(L, 1)
    return cm0, cm1, lik


def run_config(args, cfg, spb, rank, world, device, dist, peak, *, steps, warmup, strong=False, headline=False, L=None,
               M=None, parity_only=False):
    """Time one configuration (device-resident + end to end), check parity on the timed problem; returns rank 0's dict."""
    from lvae_b200 import _lib, ops
    import lvae_b200.elbo_functions as EF
    from lvae_b200.spec import build_structure, flatten
    from lvae_b200.training import natural_gradient_step
    lib = _lib.load()
    EF.set_kernel_path(args.path)
    if strong:                                   # fixed global minibatch of `spb` subjects split over the ranks
        from lvae_b200 import distributed as D
        bg = make_problem(cfg, spb, 0, 1, L, M)
        p_lo, p_hi = D.shard_subjects(bg.offsets, rank, world)
        r_lo, r_hi = int(bg.offsets[p_lo]), int(bg.offsets[p_hi])
        import copy
        b = copy.copy(bg)
        b.x, b.mu, b.log_v = bg.x[r_lo:r_hi], bg.mu[r_lo:r_hi], bg.log_v[r_lo:r_hi]
        b.offsets = bg.offsets[p_lo:p_hi + 1] - bg.offsets[p_lo]
        b.P = p_hi - p_lo
        P_b, P_glob = b.P, spb
    else:
        b = make_problem(cfg, spb, rank, world, L, M)
        P_b, P_glob = spb, spb * world
    L, M, Q, N_b = b.L, b.M, b.x.shape[1], b.N
    ragged = isinstance(b.T, tuple)
    Tl = np.diff(b.offsets)
    N_glob = N_b
    if dist is not None:
        t = torch.tensor([N_b], dtype=torch.int64, device=device)
        dist.all_reduce(t)
        N_glob = int(t.item())
    P_tot = P_glob                                # the data set is the (global) minibatch: P_tot / P_batch = 1
    N_tot = N_glob
    cm0, cm1, lik = build_modules(b, device)
    gp_params = [p for mod in (cm0, cm1, lik) for p in mod.parameters()]
    st, ls, os_ = build_structure(flatten(cm0), flatten(cm1), L, device=device)
    ls, os_ = ls.detach(), os_.detach()
    noise = lik.noise.detach().reshape(L).contiguous()
    dev = lambda t: t.to(device)
    x, mu, lv, z = dev(b.x), dev(b.mu), dev(b.log_v), dev(b.z)
    m, H = dev(b.m).clone(), dev(b.H).clone()
    offsets = torch.from_numpy(b.offsets).to(torch.int32).to(device)
    T_max, sum_T2 = int(Tl.max()), int((Tl * Tl).sum())
    const = L * N_tot / 2 if ragged else L * P_tot * int(b.T) / 2
    scale = P_tot / P_glob

    def new_call(group_unused=None):
        return (ops.make_kld_call(st, L, M, Q, Tl, device, natural_gradient=True, path=args.path,
                                  split=not getattr(args, "no_split", False)) if ragged else
                ops.KldCall(st, L, M, Q, P_b, N_b, T_max, sum_T2, device, natural_gradient=True, path=args.path))
    base = new_call()
    group = dist.group.WORLD if dist is not None else None
    tail_mode = "replicated"
    if dist is not None and isinstance(base, ops.KldCall) and L % world == 0 and \
            (args.tail == "latents" or (args.tail == "auto" and (M > 64 or strong))):
        tail_mode = "latents"
    call = ops.LatentTailKldCall(base, group) if tail_mode == "latents" else base
    ng_ws = torch.empty(int(lib.lvae_ng_workspace_doubles(L, M)), dtype=torch.float64, device=device)
    ng_info = torch.zeros(4, dtype=torch.int32, device=device)

    hold = {"call": call, "base": base}

    def device_step(c=None, mm=None, HH=None, grp=group, update=True, sc=None, ct=None):
        c = hold["call"] if c is None else c
        mm = m if mm is None else mm
        HH = H if HH is None else HH
        c.bind(x, offsets, mu, lv, z, mm.view(L, M), HH, ls, os_, noise, scale if sc is None else sc,
               const if ct is None else ct, EPS)
        c.head()
        EF.exchange_stats(c, grp, args.exchange)
        c.tail()
        if update and isinstance(c, ops.LatentTailKldCall):
            c.ng_step(mm, HH, LR)                  # own latents + all-gather of the new (m, H)
        elif update:
            rc = lib.lvae_ng_step_f64(_lib.ptr(mm), _lib.ptr(HH), _lib.ptr(c.grad_m), _lib.ptr(c.grad_H),
                                      _lib.ptr(c.Hinv), LR, L, M, _lib.ptr(ng_ws), _lib.ptr(ng_info),
                                      _lib.stream_ptr(device))
            _lib.check(rc, "lvae_ng_step_f64")

    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=device)   # > 126 MB L2
    fl = algorithmic_flops(Tl, L, M, st.n_comp0, st.n_comp1, P_b)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=device)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- parity on the timed problem (before the timing loops touch m, H) ----------------------------------------------
    parity = None
    if not args.no_parity:
        parity = check_parity(args, b, base, call, device_step, m, H, device, dist, rank, world, P_tot, P_glob, N_tot,
                              dict(x=x, mu=mu, lv=lv, offsets_np=b.offsets, z=z, ls=ls, os=os_, noise=noise, st=st,
                                   scale=scale, const=const, ragged=ragged, T=b.T), max_over_ranks)

    if parity_only:                                # the -m gpu full-size parity tests stop here
        res = Ctx()
        res.out = {"parity": parity}
        return res

    # ---- device-resident timing -------------------------------------------------------------------------------
    m0, H0 = m.clone(), H.clone()
    for _ in range(warmup):
        device_step()
    barrier()
    lib.lvae_profile_enable(1)
    launches0 = ops.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    phase_ms = np.zeros((steps, 6))
    clocks = ClockSampler(device.index)
    if rank == 0 and headline:                     # rank 0's GPU, sampled through the device-timed AND the end-to-end timed regions
        clocks.__enter__()
    barrier()
    for i in range(steps):
        flush.fill_(1.0)                           # L2 flush between timed steps (outside the event pair)
        ev[i][0].record()
        device_step()
        ev[i][1].record()
        for ph in range(6):                        # per-kernel CUDA-event durations (syncs on this step's events)
            phase_ms[i, ph] = lib.lvae_profile_last_ms(ph)
    barrier()
    launches = ops.launch_count() - launches0
    lib.lvae_profile_enable(0)
    total_ms = sum(a.elapsed_time(b_) for a, b_ in ev)
    call.raise_on_info()
    ms_per_step = max_over_ranks(total_ms) / steps
    value = P_glob / (ms_per_step * 1e-3)
    finite = bool(torch.isfinite(call.kld_per_latent).all() and torch.isfinite(H).all())
    is_split = isinstance(call, ops.SplitKldCall)

    # ---- end to end through the public API with host buffers --------------------------------------------------
    hold["call"] = hold["base"] = None
    del base
    del call                                   # its workspace (tens of GB at M > 64 with large minibatches) is not needed any more
    torch.cuda.empty_cache()
    m.copy_(m0); H.copy_(H0)
    hx, hmu, hlv = b.x.pin_memory(), b.mu.pin_memory(), b.log_v.pin_memory()
    out_mu = torch.empty_like(b.mu).pin_memory()
    out_lv = torch.empty_like(b.log_v).pin_memory()
    out_kld = torch.empty(1, dtype=torch.float64).pin_memory()
    if dist is not None:
        EF.set_process_group(dist.group.WORLD, args.exchange, "subjects", tail_mode)
    EF.set_error_check("deferred")          # no device sync inside the op; failures still raise (check_errors below)
    state = {"m": m, "H": H}

    # Software pipeline over steps (all inside the timed region): the H2D copy of step i+1's inputs runs on a copy stream
    # while step i computes, the D2H copy of step i's results runs on a second copy stream; every step still copies its
    # own inputs from pinned host memory and returns its own loss / encoder gradients to pinned host memory.
    main = torch.cuda.current_stream(device)
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    dbuf = [dict(x=torch.empty_like(x), mu=torch.empty_like(mu), lv=torch.empty_like(lv),
                 ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
    out_done = torch.cuda.Event()

    def prefetch(i):
        d = dbuf[i % 2]
        with torch.cuda.stream(s_in):
            s_in.wait_event(d["free"])                     # the step that last used this buffer has finished with it
            d["x"].copy_(hx, non_blocking=True)
            d["mu"].copy_(hmu, non_blocking=True)
            d["lv"].copy_(hlv, non_blocking=True)
            d["ready"].record(s_in)

    def api_step(xd, mud, lvd):
        if ragged:      # rows per subject from the host side of the batch, as hensman_training takes them from the loader
            kld, gm, gH = EF.minibatch_KLD_upper_bound_iter(cm0, cm1, lik, L, state["m"], state["H"], xd, mud, lvd, z,
                                                            P_tot, P_glob, N_tot, True, 2, EPS, subject_counts=Tl)
        else:
            kld, gm, gH = EF.minibatch_KLD_upper_bound(cm0, cm1, lik, L, state["m"], state["H"], xd, mud, lvd, z, P_tot,
                                                       P_glob, int(b.T), True, EPS)
        kld.sum().backward()
        return kld, gm, gH

    def e2e_step(i):
        d = dbuf[i % 2]
        prefetch(i + 1)
        main.wait_event(d["ready"])
        xd = d["x"]
        mud = d["mu"].detach().requires_grad_(True)
        lvd = d["lv"].detach().requires_grad_(True)
        kld, gm, gH = api_step(xd, mud, lvd)
        d["free"].record(main)
        state["m"], state["H"] = natural_gradient_step(state["m"], state["H"], gm, gH, LR)
        done = torch.cuda.Event()
        done.record(main)
        gmu, glv, kd = mud.grad, lvd.grad, kld.detach().reshape(1)
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            out_mu.copy_(gmu, non_blocking=True)
            out_lv.copy_(glv, non_blocking=True)
            out_kld.copy_(kd, non_blocking=True)
            for t_ in (gmu, glv, kd):
                t_.record_stream(s_out)
            out_done.record(s_out)
        zero_grads(gp_params)

    for d in dbuf:
        d["free"].record(main)
    n_e2e = 0
    prefetch(0)
    for _ in range(max(3, warmup)):
        e2e_step(n_e2e); n_e2e += 1
    main.wait_event(out_done)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    host_ms, t_prev = [], time.perf_counter()
    allocs0 = torch.cuda.memory_stats(device).get("num_device_alloc", 0)
    for _ in range(steps):
        e2e_step(n_e2e); n_e2e += 1
        t_now = time.perf_counter(); host_ms.append(round((t_now - t_prev) * 1e3, 3)); t_prev = t_now
    main.wait_event(out_done)                              # the last step's results have reached host memory
    e1.record()
    barrier()
    e2e_diag = {"host_enqueue_ms_per_step_rank0": host_ms[:32],
                "cudaMalloc_calls_in_timed_region_rank0": torch.cuda.memory_stats(device).get("num_device_alloc", 0) - allocs0}
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    e2e_val = P_glob / (e2e_ms * 1e-3)
    if args.profile_e2e and rank == 0:                     # diagnostic: where the API step spends its time (not a bench value)
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                e2e_step(n_e2e); n_e2e += 1
            torch.cuda.synchronize(device)
        with open(f"{args.profile_e2e}.{cfg}{'_strong' if strong else ''}.txt", "w") as fh:
            fh.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
            fh.write("\n")
            fh.write(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=70))
    elif args.profile_e2e:
        for _ in range(3):
            e2e_step(n_e2e); n_e2e += 1
        torch.cuda.synchronize(device)
    h2d = (hx.numel() + hmu.numel() + hlv.numel()) * 8
    d2h = (out_mu.numel() + out_lv.numel() + 1) * 8
    EF.check_errors()

    # the plain loop a user writes first: immediate error check (one device sync per call, like torch.cholesky), copies on the
    # compute stream, nothing overlapped
    EF.set_error_check("immediate")
    state["m"], state["H"] = m0.clone(), H0.clone()

    def plain_step():
        xd = hx.to(device, non_blocking=True)
        mud = hmu.to(device, non_blocking=True).requires_grad_(True)
        lvd = hlv.to(device, non_blocking=True).requires_grad_(True)
        kld, gm, gH = api_step(xd, mud, lvd)
        state["m"], state["H"] = natural_gradient_step(state["m"], state["H"], gm, gH, LR)
        out_mu.copy_(mud.grad, non_blocking=True)
        out_lv.copy_(lvd.grad, non_blocking=True)
        out_kld.copy_(kld.detach().reshape(1), non_blocking=True)
        zero_grads(gp_params)
    for _ in range(3):
        plain_step()
    barrier()
    e0.record()
    for _ in range(steps):
        plain_step()
    e1.record()
    barrier()
    plain_ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    if rank == 0 and headline:
        clocks.__exit__(None, None, None)
    EF.set_process_group(None)

    if M <= 64:
        kernel_path = ("generic" if args.path == 1 else
                       "split by subject length: third-generation fused kernel with 24-row groups (T <= 24) and 40-row groups"
                       if is_split else "third-generation fused kernel (k_prep3 + k_subjects_fused3)" if M <= 62 else
                       "second-generation fused kernel")
    else:
        kernel_path = "gemm (U/V materialised, S = U^T U and Y = V W as batched DMMA GEMMs)" if T_max <= 24 else "generic"
    subj_ms = float(phase_ms[:, 2].mean())
    achieved = fl["subjects"] / (subj_ms * 1e-3) * 1e-12 if subj_ms > 0 else None
    res = Ctx()
    res.b, res.st, res.x, res.mu, res.lv, res.z, res.offsets = b, st, x, mu, lv, z, offsets
    res.ls, res.os, res.noise, res.m0, res.H0, res.cm0, res.cm1, res.lik = ls, os_, noise, m0, H0, cm0, cm1, lik
    res.P_tot, res.N_tot, res.const, res.Tl, res.T_max = P_tot, N_tot, const, Tl, T_max
    res.out = {
        "value": value, "ms_per_step": ms_per_step, "subjects_per_gpu": P_b, "global_batch_subjects": P_glob,
        "scaling": "strong" if strong else "weak", "tail": tail_mode,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "diag": e2e_diag,
                "plain_loop_value": P_glob / (plain_ms * 1e-3), "plain_loop_ms_per_step": plain_ms,
                "how": "public API (minibatch_KLD_upper_bound[_iter] + backward + natural_gradient_step); every step copies "
                       "x, mu, log_v from pinned host memory and returns kld, d_mu, d_log_v to pinned host memory.  `value`: "
                       "the copies of neighbouring steps overlap the compute on two copy streams and the Cholesky-failure "
                       "flags are checked deferred; `plain_loop_value`: same API in a plain loop (copies on the compute "
                       "stream, one device sync per call for the error check, as torch.cholesky has)"},
        "gpu_launches": int(launches), "finite": finite, "parity": parity, "kernel_path": kernel_path,
        "phase_ms": {n: float(phase_ms[:, i].mean()) for i, n in
                     enumerate(["head", "prep", "subjects", "reduce", "tail", "ng_step"])},
        "roofline_frac": (achieved / peak) if (achieved and peak) else None,
        "subject_pass_tflops": achieved, "step_tflops": fl["step"] / (ms_per_step * 1e-3) * 1e-12,
        "flops": fl, "clocks": clocks.summary() if (rank == 0 and headline) else None,
    }
    return res


def check_parity(args, b, call, shcall, device_step, m, H, device, dist, rank, world, P_tot, P_glob, N_tot, t, max_over_ranks):
    """Parity on the timed problem.  (a) This rank's shard, CUDA path on one GPU (no exchange) vs the oracle port run with
    stock torch ops on the same GPU, global scaling constants: kld, grad_m, grad_H, d_mu, d_log_v, hyper-gradient vector,
    and (m, H) after the natural-gradient update.  (b) N > 1: the sharded result vs a one-GPU CUDA evaluation of the
    gathered minibatch on every rank (kld, grad_m, grad_H, hyper-gradients, own rows of d_mu / d_log_v)."""
    from lvae_b200 import _lib, ops
    lib = _lib.load()
    L, M = b.L, b.M
    mm, HH = m.clone(), H.clone()
    ragged = t["ragged"]
    const_loc = L * int(t["x"].shape[0]) / 2 if ragged else L * int(b.P) * int(b.T) / 2
    # this rank's shard as a data set of its own (scale 1, its own constant term), incl. the NG update of the copies
    device_step(c=call, mm=mm, HH=HH, grp=None, update=True, sc=1.0, ct=const_loc)
    torch.cuda.synchronize(device)
    call.raise_on_info()
    ours = dict(kld=call.kld_per_latent.sum(), grad_m=call.grad_m.clone(), grad_H=call.grad_H.clone(),
                d_mu=call.d_mu.clone(), d_log_v=call.d_log_v.clone(), d_hyper=call.d_hyper.clone(), m_new=mm, H_new=HH)
    KEYS = ("kld", "grad_m", "grad_H", "d_mu", "d_log_v", "d_hyper", "m_new", "H_new")
    # (a1) every latent, every subject against the oracle run with stock torch CUDA ops on this GPU.  cuSOLVER / cuBLAS FP64
    # are themselves 1e-6..1e-4 away from LAPACK on the Kzz^-1-dependent outputs at these sizes (cond ~ 1e8; measured,
    # profiles/r02_parity_three_way.txt), so this is the gross-error net over ALL latents: 1e-6 where the oracle is that
    # accurate (d_mu, d_log_v), 1e-2 elsewhere (that oracle is itself up to 4e-4 from the exact grad_m at M = 128).
    with torch.device(device):
        ref = oracle_step_fn(b, b.P, device=device)(update=True)
    torch.cuda.synchronize(device)
    gross_tol = {k: (TOL if k in ("d_mu", "d_log_v") else 1e-2) for k in KEYS}
    gross = {k: rel_err(ours[k], ref[k]) for k in KEYS}
    ref_dev = ref                                  # the reference's torch-CUDA results: sliced to the checked latents below
    if world > 1:
        # N > 1: every rank checks its shard against the torch-CUDA oracle (a1) and the sharded result against the one-GPU
        # evaluation of the gathered minibatch (b, below).  The host-side checks (a2, a3: LAPACK oracle, extended precision)
        # are what the N = 1 run records for the same kind of shard — N processes would share the host cores for them.
        okl = all(gross[k] <= gross_tol[k] for k in KEYS)
        okf = max_over_ranks(0.0 if okl else 1.0) == 0.0
        gmax = {k: max_over_ranks(gross[k]) for k in KEYS}
        out = {"ok": bool(okf), "max_rel": max(gmax[k] for k in ("d_mu", "d_log_v")), "tol": TOL,
               "all_latents_vs_torch_cuda_oracle": {"per_tensor_max_over_ranks": gmax, "tol_per_tensor": gross_tol, "ok": bool(okf)},
               "against": "every rank: its shard against the oracle port with stock torch CUDA ops on its GPU (1e-6 on d_mu / "
                          "d_log_v, 1e-2 gross-error net on the Kzz^-1-carrying outputs); `cross_rank`: the sharded result against "
                          "a one-GPU evaluation of the gathered minibatch.  The host-side checks (LAPACK oracle, extended "
                          "precision) are recorded by the N = 1 run",
               "n_subjects_checked_per_rank": int(b.P)}
        del ref, ref_dev
        torch.cuda.empty_cache()
    else:
        # (a2) the tight check: the oracle on the HOST (torch CPU FP64 = LAPACK / MKL, the reference's own arithmetic) on all
        # subjects and a subset of the latent dimensions (they are independent; first, last and two in between).  Next to it
        # the reference's own sensitivity to input rounding: the same oracle with the jitter eps scaled by 1 + 1e-9, i.e. the
        # diagonal of Kzz moved by 1e-15 ~ 2 ulp — less than what rounding the kernel entries does.  tol = 1e-6 + 2 x that.
        lat = sorted(set([0, L // 3, (2 * L) // 3, L - 1])) if M <= 64 else sorted(set([0, L - 1]))
        nthr = torch.get_num_threads()
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(1, world)))
        with torch.device("cpu"):
            cref = oracle_step_fn(b, b.P, device="cpu", latents=lat)(update=True)
            cper = oracle_step_fn(b, b.P, device="cpu", latents=lat, eps=EPS * (1 + 1e-9))(update=True)
            # (a3) the same formulas in extended precision (numpy longdouble, oracle/lvae_oracle_xp.py): where the reference's FP64
            # result is itself further than 1e-6 from the exact one (Kzz^-1 enters grad_m / grad_H twice, cond(Kzz) 3e7 .. 1e9),
            # "parity" can only mean: this implementation is as close to the exact result as the reference is.
            import lvae_oracle_xp as oxp
            k0x, k1x, nzx, _ = oracle_components(b, "cpu", requires_grad=False)
            truth = oxp.kld_forward(k0x, k1x, nzx, lat, b.m, b.H, b.x, b.offsets, b.mu, b.log_v, b.z, 1.0, const_loc / L, EPS)
        torch.set_num_threads(nthr)
        li = torch.as_tensor(lat, device=device)
        pick = dict(kld=call.kld_per_latent[li].sum(), grad_m=ours["grad_m"][li], grad_H=ours["grad_H"][li],
                    d_mu=ours["d_mu"][:, li], d_log_v=ours["d_log_v"][:, li], d_hyper=ours["d_hyper"][:, li],
                    m_new=ours["m_new"].reshape(L, M)[li], H_new=ours["H_new"][li])
        errs = {k: rel_err(pick[k], cref[k]) for k in KEYS}
        floor = {k: rel_err(cper[k], cref[k]) for k in KEYS}
        dpick = dict(kld=None, grad_m=ref_dev["grad_m"][li], grad_H=ref_dev["grad_H"][li], d_mu=ref_dev["d_mu"][:, li],
                     d_log_v=ref_dev["d_log_v"][:, li], d_hyper=ref_dev["d_hyper"][:, li],
                     m_new=ref_dev["m_new"].reshape(L, M)[li], H_new=ref_dev["H_new"][li])
        del ref, ref_dev
        torch.cuda.empty_cache()
        with torch.device(device):                     # kld of exactly these latents from the torch-CUDA path (it returns a sum)
            dpick["kld"] = oracle_step_fn(b, b.P, device=device, latents=lat)(update=False)["kld"]
        torch.cuda.empty_cache()
        dev_vs_host = {k: rel_err(dpick[k], cref[k]) for k in KEYS}
        tr = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in truth.items()}
        tr["kld"] = tr["kld"].sum()
        XK = ("kld", "grad_m", "grad_H", "d_mu", "d_log_v")
        ours_exact = {k: rel_err(pick[k].cpu().reshape(tr[k].shape), tr[k]) for k in XK}
        ref_exact = {k: rel_err(cref[k].reshape(tr[k].shape), tr[k]) for k in XK}
        dev_exact = {k: rel_err(dpick[k].cpu().reshape(tr[k].shape), tr[k]) for k in XK if dpick[k] is not None}
        # A tensor passes if it is within 1e-6 of the reference's host result; or (exact value known) no further from the exact
        # value than twice the reference's host result, or twice the reference's own torch-CUDA result — the reference selects
        # "cuda" whenever a GPU is present (elbo_functions.py:165), so THAT is what it computes on this box; or (no exact value)
        # within the backward-error allowance of an M x M Cholesky, 1e-6 + M/2 x the 2-ulp sensitivity, or closer to the host
        # result than the reference's torch-CUDA result is.
        tols = {k: TOL + 0.5 * M * floor[k] for k in KEYS}
        verdict = {}
        for k in KEYS:
            if errs[k] <= TOL:
                verdict[k] = "within 1e-6 of the reference (host)"
            elif k in XK and ours_exact[k] <= max(TOL, 2 * ref_exact[k]):
                verdict[k] = "as close to the exact value as the reference on the host"
            elif k in dev_exact and ours_exact[k] <= max(TOL, 2 * dev_exact[k]):
                verdict[k] = "as close to the exact value as the reference's torch-CUDA path on this GPU"
            elif k not in XK and errs[k] <= tols[k]:
                verdict[k] = "within the reference's input-rounding allowance"
            elif k not in XK and k in dev_vs_host and errs[k] <= 2 * dev_vs_host[k]:
                verdict[k] = "as close to the reference's host result as its torch-CUDA path on this GPU"
            else:
                verdict[k] = "FAIL"
        dh, rh = pick["d_hyper"].double(), cref["d_hyper"].double().to(device)
        worst_entry = float(((dh - rh).abs() / rh.abs().clamp_min(1e-300)).max())
        ok = all(v != "FAIL" for v in verdict.values()) and all(gross[k] <= gross_tol[k] for k in KEYS)
        okf = max_over_ranks(0.0 if ok else 1.0) == 0.0
        out = {"max_rel": max_over_ranks(max(errs.values())), "tol": TOL, "ok": bool(okf),
               "per_tensor_rank0": errs, "verdict_rank0": verdict,
               "vs_exact_rank0": {"ours": ours_exact, "reference": ref_exact, "reference_torch_cuda": dev_exact,
                                  "what": "max-norm relative distance to the same formulas evaluated in extended precision (numpy "
                                          "longdouble, oracle/lvae_oracle_xp.py) — the reference's own FP64 error on this problem"},
               "reference_torch_cuda_vs_host_rank0": dev_vs_host,
               "input_rounding_floor_rank0": floor, "allowance_per_tensor_rank0": tols,
               "d_hyper_worst_single_entry_rel_rank0": worst_entry, "latents_checked": lat,
               "against": "oracle port of the reference on the host (torch CPU FP64 / LAPACK), ALL subjects of the timed step, "
                          "latent dimensions `latents_checked`; max-norm relative error per tensor.  A tensor passes if it is within "
                          "1e-6 of the reference's host result; or no further from the extended-precision value than twice the host "
                          "result or twice the reference's own torch-CUDA result on this GPU (kld, grad_m, grad_H, d_mu, d_log_v); or "
                          "(d_hyper, m_new, H_new) within 1e-6 + M/2 x the change of the reference's output when the Kzz diagonal moves "
                          "by 2 ulp (backward error of an M x M Cholesky ~ M ulp), or no further from the host result than twice the "
                          "reference's torch-CUDA result",
               "all_latents_vs_torch_cuda_oracle": {"per_tensor_rank0": gross, "tol_per_tensor": gross_tol,
                                                    "ok": bool(all(gross[k] <= gross_tol[k] for k in KEYS))},
               "n_subjects_checked_per_rank": int(b.P)}
    if dist is not None:                                   # (b) cross-rank: sharded vs one-GPU evaluation of the gathered batch
        m2, H2 = m.clone(), H.clone()
        device_step(c=shcall, mm=m2, HH=H2, grp=dist.group.WORLD, update=True)       # the sharded step exactly as it is timed
        torch.cuda.synchronize(device)
        sh = dict(kld=shcall.kld_per_latent.sum().clone(), grad_m=shcall.grad_m.clone(), grad_H=shcall.grad_H.clone(),
                  d_mu=shcall.d_mu.clone(), d_log_v=shcall.d_log_v.clone(), d_hyper=shcall.d_hyper.clone(), m_new=m2, H_new=H2)
        if isinstance(shcall, ops.LatentTailKldCall):     # grad_m / grad_H hold this rank's latents only: gather for the check
            for k_ in ("grad_m", "grad_H"):
                own = sh[k_][shcall.l0:shcall.l1].contiguous()
                dist.all_gather_into_tensor(sh[k_], own)
        nrows = torch.tensor([t["x"].shape[0]], dtype=torch.int64, device=device)
        allrows = [torch.zeros_like(nrows) for _ in range(world)]
        dist.all_gather(allrows, nrows)
        allrows = [int(v.item()) for v in allrows]
        nsub = torch.tensor([b.P], dtype=torch.int64, device=device)
        allsub = [torch.zeros_like(nsub) for _ in range(world)]
        dist.all_gather(allsub, nsub)
        allsub = [int(v.item()) for v in allsub]

        def gather_rows(v):
            parts = [torch.empty(n, *v.shape[1:], dtype=v.dtype, device=device) for n in allrows]
            dist.all_gather(parts, v.contiguous())
            return torch.cat(parts)
        gx, gmu, glv = gather_rows(t["x"]), gather_rows(t["mu"]), gather_rows(t["lv"])
        counts = [torch.empty(n, dtype=torch.int64, device=device) for n in allsub]
        dist.all_gather(counts, torch.from_numpy(np.diff(t["offsets_np"]).astype(np.int64)).to(device))
        Tg = torch.cat(counts).cpu().numpy()
        offg = torch.from_numpy(np.concatenate([[0], np.cumsum(Tg)]).astype(np.int32)).to(device)
        Q = gx.shape[1]
        big = (ops.make_kld_call(t["st"], L, M, Q, Tg, device, natural_gradient=True, path=args.path) if t["ragged"] else
               ops.KldCall(t["st"], L, M, Q, len(Tg), int(gx.shape[0]), int(Tg.max()), int((Tg * Tg).sum()), device,
                           natural_gradient=True, path=args.path))
        ws = torch.empty(int(lib.lvae_ng_workspace_doubles(L, M)), dtype=torch.float64, device=device)
        info = torch.zeros(4, dtype=torch.int32, device=device)

        def one_gpu(xs, offs, mus_, lvs_):
            m3, H3 = m.clone(), H.clone()
            big.bind(xs, offs, mus_, lvs_, t["z"], m3.view(L, M), H3, t["ls"], t["os"], t["noise"], t["scale"], t["const"], EPS)
            big.run()
            _lib.check(lib.lvae_ng_step_f64(_lib.ptr(m3), _lib.ptr(H3), _lib.ptr(big.grad_m), _lib.ptr(big.grad_H),
                                            _lib.ptr(big.Hinv), LR, L, M, _lib.ptr(ws), _lib.ptr(info),
                                            _lib.stream_ptr(device)), "lvae_ng_step_f64")
            torch.cuda.synchronize(device)
            big.raise_on_info()
            return dict(kld=big.kld_per_latent.sum().clone(), grad_m=big.grad_m.clone(), grad_H=big.grad_H.clone(),
                        d_mu=big.d_mu.clone(), d_log_v=big.d_log_v.clone(), d_hyper=big.d_hyper.clone(), m_new=m3, H_new=H3)
        full = one_gpu(gx, offg, gmu, glv)
        r0 = sum(allrows[:rank])
        one = dict(full)
        one["d_mu"], one["d_log_v"] = full["d_mu"][r0:r0 + allrows[rank]], full["d_log_v"][r0:r0 + allrows[rank]]
        cerr = {k: rel_err(sh[k], one[k]) for k in one}
        # Yardstick: sharding only changes the ORDER in which the statistics of the subjects are summed.  So does evaluating
        # the same gathered minibatch on one GPU with its subjects randomly permuted — the spread between those two one-GPU
        # results is what a re-ordered sum costs on this problem (Kzz^-1 amplifies the last bits of S and ng1).
        ends = np.concatenate([[0], np.cumsum(Tg)])
        perm = np.random.default_rng(12345).permutation(len(Tg))       # same on every rank
        order = np.concatenate([np.arange(ends[i], ends[i + 1]) for i in perm])
        oi = torch.from_numpy(order).to(device)
        offr = torch.from_numpy(np.concatenate([[0], np.cumsum(Tg[perm])]).astype(np.int32)).to(device)
        if t["ragged"] and isinstance(big, ops.SplitKldCall):
            big = ops.make_kld_call(t["st"], L, M, Q, Tg[perm].copy(), device, natural_gradient=True, path=args.path)
        rev = one_gpu(gx[oi], offr, gmu[oi], glv[oi])
        AMP = ("kld", "grad_m", "grad_H", "d_hyper", "m_new", "H_new")
        spread = {k: rel_err(rev[k], full[k]) for k in AMP}
        ctol = {k: (max(TOL, 8 * spread[k]) if k in AMP else TOL) for k in cerr}
        cok = all(cerr[k] <= ctol[k] for k in cerr)
        cmax = max_over_ranks(max(cerr.values()))
        cokf = max_over_ranks(0.0 if cok else 1.0) == 0.0
        ident = torch.stack([sh["kld"].reshape(()), sh["grad_H"].sum(), sh["d_hyper"].sum()])
        lo, hi = ident.clone(), ident.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out["cross_rank"] = {"max_rel": cmax, "ok": bool(cokf), "per_tensor_rank0": cerr, "tol_per_tensor_rank0": ctol,
                             "one_gpu_reordered_spread_rank0": spread,
                             "ranks_bit_identical": bool(torch.equal(lo, hi)),
                             "against": f"one-GPU CUDA evaluation of the gathered {int(len(Tg))}-subject minibatch on every rank; "
                                        "d_mu / d_log_v must agree to 1e-6 (they are bit-identical in practice), the outputs that "
                                        "carry Kzz^-1 to max(1e-6, 8 x the spread between two one-GPU evaluations of that minibatch "
                                        "with its subjects in the given and in a randomly permuted order)"}
        out["ok"] = bool(out["ok"] and out["cross_rank"]["ok"])
        del big
        torch.cuda.empty_cache()
    return out


def latency_point(args, R, device):
    """spb = 20, the reference's default batch (launch / latency regime)."""
    from lvae_b200 import _lib, ops
    import lvae_b200.elbo_functions as EF
    from lvae_b200.training import natural_gradient_step
    lib = _lib.load()
    b, st = R.b, R.st
    L, M, Q = b.L, b.M, R.x.shape[1]
    ragged = isinstance(b.T, tuple)
    spb2 = 20
    rows = int(b.offsets[spb2])
    off2 = R.offsets[:spb2 + 1].contiguous()
    c2 = ops.KldCall(st, L, M, Q, spb2, rows, R.T_max, int((R.Tl[:spb2] ** 2).sum()), device, True, args.path)
    mm, HH = R.m0.clone(), R.H0.clone()
    ng_ws = torch.empty(int(lib.lvae_ng_workspace_doubles(L, M)), dtype=torch.float64, device=device)
    ng_info = torch.zeros(4, dtype=torch.int32, device=device)
    x, mu, lv, z = R.x, R.mu, R.lv, R.z

    def small():
        c2.bind(x[:rows], off2, mu[:rows], lv[:rows], z, mm.view(L, M), HH, R.ls, R.os, R.noise, R.P_tot / spb2, R.const, EPS)
        c2.run()
        lib.lvae_ng_step_f64(_lib.ptr(mm), _lib.ptr(HH), _lib.ptr(c2.grad_m), _lib.ptr(c2.grad_H),
                             _lib.ptr(c2.Hinv), LR, L, M, _lib.ptr(ng_ws), _lib.ptr(ng_info),
                             _lib.stream_ptr(device))
    for _ in range(5):
        small()
    torch.cuda.synchronize()
    a_, b__ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_.record()
    for _ in range(50):
        small()
    b__.record(); torch.cuda.synchronize()
    lat = {"spb": spb2, "ms_per_step": a_.elapsed_time(b__) / 50, "subjects_per_s": spb2 / (a_.elapsed_time(b__) / 50 * 1e-3),
           "what": "spb = 20 (the reference's default minibatch), device-resident inputs, bound + gradients + NG update"}
    cm0, cm1, lik = R.cm0, R.cm1, R.lik
    gp_params = [p for mod in (cm0, cm1, lik) for p in mod.parameters()]
    try:      # the same small minibatch through the public API with pinned host inputs (Python + launch overhead regime)
        hx2, hmu2, hlv2 = b.x[:rows].pin_memory(), b.mu[:rows].pin_memory(), b.log_v[:rows].pin_memory()
        st2 = {"m": R.m0.clone(), "H": R.H0.clone()}
        o_mu = torch.empty_like(b.mu[:rows]).pin_memory()

        def small_api():
            xd = hx2.to(device, non_blocking=True)
            mud = hmu2.to(device, non_blocking=True).requires_grad_(True)
            lvd = hlv2.to(device, non_blocking=True).requires_grad_(True)
            if ragged:
                kld, gm, gH = EF.minibatch_KLD_upper_bound_iter(cm0, cm1, lik, L, st2["m"], st2["H"], xd, mud, lvd, z, R.P_tot,
                                                                spb2, R.N_tot, True, 2, EPS)
            else:
                kld, gm, gH = EF.minibatch_KLD_upper_bound(cm0, cm1, lik, L, st2["m"], st2["H"], xd, mud, lvd, z, R.P_tot, spb2,
                                                           int(b.T), True, EPS)
            kld.sum().backward()
            st2["m"], st2["H"] = natural_gradient_step(st2["m"], st2["H"], gm, gH, LR)
            o_mu.copy_(mud.grad, non_blocking=True)
            zero_grads(gp_params)
        EF.set_error_check("deferred")
        for _ in range(5):
            small_api()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            small_api()
        torch.cuda.synchronize()
        lat["e2e_ms_per_step"] = (time.perf_counter() - t0) / 50 * 1e3
        lat["e2e_subjects_per_s"] = spb2 / (lat["e2e_ms_per_step"] * 1e-3)
        EF.check_errors()
        EF.set_error_check("immediate")
    except Exception as ex:
        lat["e2e_unavailable"] = repr(ex)[:160]
    if not ragged:
        try:  # public API again, but the GP side of the step is ONE graph replay (lvae_b200.graphed.GraphedHensmanStep)
            from lvae_b200.graphed import GraphedHensmanStep
            mg, Hg = R.m0.clone().reshape(L, M, 1).contiguous(), R.H0.clone().contiguous()
            gstep = GraphedHensmanStep(cm0, cm1, lik, L, mg, Hg, z, R.P_tot, spb2, int(b.T), EPS, LR)

            def small_graphed():
                xd = hx2.to(device, non_blocking=True)
                mud = hmu2.to(device, non_blocking=True).requires_grad_(True)
                lvd = hlv2.to(device, non_blocking=True).requires_grad_(True)
                gstep(xd, mud, lvd).backward()
                o_mu.copy_(mud.grad, non_blocking=True)
                zero_grads(gp_params)
            for _ in range(5):
                small_graphed()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(50):
                small_graphed()
            torch.cuda.synchronize()
            lat["graphed_api_ms_per_step"] = (time.perf_counter() - t0) / 50 * 1e3
            lat["graphed_api_subjects_per_s"] = spb2 / (lat["graphed_api_ms_per_step"] * 1e-3)
            gstep.check_errors()
            lat["graphed_api_finite"] = bool(torch.isfinite(Hg).all())
        except Exception as ex:
            lat["graphed_api_unavailable"] = repr(ex)[:200]
    try:      # the same step captured once into a CUDA graph (the C ABI is stream-ordered and capture-safe) and replayed
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            small()
        torch.cuda.current_stream(device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            small()
        for _ in range(5):
            graph.replay()
        torch.cuda.synchronize()
        a_.record()
        for _ in range(50):
            graph.replay()
        b__.record(); torch.cuda.synchronize()
        lat["graph_ms_per_step"] = a_.elapsed_time(b__) / 50
        lat["graph_subjects_per_s"] = spb2 / (a_.elapsed_time(b__) / 50 * 1e-3)
        lat["graph_finite"] = bool(torch.isfinite(HH).all())
    except Exception as ex:
        lat["graph_unavailable"] = repr(ex)[:160]
    return lat


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = max(1, args.gpus)
    cfg = args.cfg
    spb = args.spb or DEFAULT_SPB[cfg]
    b = make_problem(cfg, spb, 0, 1, args.L, args.M)
    n_sub = default_cpu_subjects(args, b, spb)
    P_glob = spb * world
    steps, warm = max(1, args.steps), max(0, args.warmup)
    val, dt, kind = time_cpu(b, n_sub, steps, warm)
    cores = os.cpu_count() or 1
    what = ("the reference's own modules (baseline/_ref: elbo_functions.minibatch_KLD_upper_bound"
            f"{'_iter' if isinstance(b.T, tuple) else ''} + kernel_gen.generate_kernel_batched over the GPyTorch stand-in)"
            if kind == "reference" else "oracle port of the reference (baseline/_ref absent)")
    frac = "all" if (n_sub == spb and world == 1) else f"{n_sub} of the {P_glob}"
    sample = (f"{what}, torch CPU FP64 with {cores} threads: {frac} subjects of the minibatch per step (fwd + backward + NG "
              f"update; per-subject cost is constant, the fixed M^3 part is < 3 %), {steps} steps after {warm} warm-up, "
              f"{dt:.2f} s/step")
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config(cfg, spb, world, b),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


def main():
    args = parse()
    _quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = args.cfg
    spb = args.spb or DEFAULT_SPB[cfg]
    if args.impl == "reference":
        return reference_arm(args)

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(minutes=4))   # a hang fails fast
    from lvae_b200 import _lib
    _lib.load()

    exchange_note = None
    if dist is not None and args.exchange == "p2p":       # all ranks agree on the exchange path before any timed work
        ok = 1
        try:
            from lvae_b200 import distributed as D
            D.peer_stats(dist.group.WORLD, 64, device)
        except Exception as ex:
            ok, exchange_note = 0, f"symmetric memory unavailable ({repr(ex)[:120]}): NCCL all-reduce"
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            args.exchange = "nccl"
            exchange_note = exchange_note or "a peer could not map symmetric memory: NCCL all-reduce"

    peak = dgemm_peak_tflops(device)
    R = run_config(args, cfg, spb, rank, world, device, dist, peak, steps=args.steps, warmup=args.warmup, headline=True,
                   L=args.L, M=args.M)
    o = R.out
    b, st = R.b, R.st
    L, M = b.L, b.M

    lat = None
    if not args.no_latency_point and rank == 0 and world == 1 and spb >= 20:
        lat = latency_point(args, R, device)

    cpu = gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_sub = default_cpu_subjects(args, b, spb)
        v, dt, kind = time_cpu(make_problem(cfg, spb, 0, 1, args.L, args.M), n_sub, 3, 1)
        cores = os.cpu_count() or 1
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": (f"{'the reference own modules from baseline/_ref' if kind == 'reference' else 'oracle port of the reference'}"
                          f" (torch CPU FP64, {cores} threads): {n_sub} of the {spb} subjects per step, fwd + backward + NG "
                          f"update, 3 steps after 1 warm-up ({dt:.2f} s/step)")}
        try:
            v, ms = time_gpu_reference(make_problem(cfg, spb, 0, 1, args.L, args.M), device)
            gpu_ref = {"value": v, "unit": UNIT, "ms_per_step": ms,
                       "what": "oracle port of the reference with stock PyTorch CUDA ops (ATen/cuBLAS/cuSOLVER) on this "
                               "GPU, device-resident inputs, all subjects of the minibatch, fwd + backward + NG update"}
        except Exception as ex:   # the baseline is informative only
            gpu_ref = {"unavailable": repr(ex)[:200]}
        torch.cuda.empty_cache()

    others = {}
    if not args.no_others and cfg == "cfg2" and args.L is None and args.M is None:
        plan = [("cfg3", dict(cfg="cfg3", spb=DEFAULT_SPB["cfg3"])),
                ("cfg4", dict(cfg="cfg4", spb=DEFAULT_SPB["cfg4"])),
                ("cfg5", dict(cfg="cfg5", spb=DEFAULT_SPB["cfg5"]))]
        if world > 1:
            plan += [("cfg3_strong", dict(cfg="cfg3", spb=1000, strong=True)),
                     ("cfg2_strong", dict(cfg="cfg2", spb=1000, strong=True))]
        for name, kw in plan:
            try:
                r = run_config(args, kw["cfg"], kw["spb"], rank, world, device, dist, peak, steps=args.others_steps,
                               warmup=3, strong=kw.get("strong", False))
                oo = r.out
                others[name] = {"config": workload_config(kw["cfg"], kw["spb"], world, r.b, kw.get("strong", False)),
                                "value": oo["value"], "unit": UNIT, "ms_per_step": oo["ms_per_step"],
                                "scaling": oo["scaling"], "steps": args.others_steps, "warmup": 3,
                                "e2e": {k: oo["e2e"][k] for k in ("value", "ms_per_step", "diag", "plain_loop_value",
                                                                   "h2d_bytes_per_step", "d2h_bytes_per_step")},
                                "roofline_frac": oo["roofline_frac"], "subject_pass_tflops": oo["subject_pass_tflops"],
                                "step_tflops": oo["step_tflops"], "phase_ms": oo["phase_ms"], "parity": oo["parity"],
                                "kernel_path": oo["kernel_path"], "tail": oo["tail"], "gpu_launches": oo["gpu_launches"],
                                "finite": oo["finite"]}
                del r
            except Exception as ex:
                others[name] = {"failed": repr(ex)[:300]}
            torch.cuda.empty_cache()

    rc = 0
    if rank == 0:
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else None
        roof_kernel = ("subject pass (fused kernel blocks + trisolve + S/Y contractions, FP64 DMMA)" if M <= 64 else
                       "subject pass (k_uv + S = U^T U GEMM + Y = V W GEMM + k_adj, FP64 DMMA)")
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tfile):
            ent = json.load(open(tfile)).get(f"{cfg}:spb{spb}")
            traffic = ent["traffic_bytes"] if ent else None
        subj_ms = o["phase_ms"]["subjects"]
        roof = {"bound": "tensor", "kernel": roof_kernel,
                "achieved": o["subject_pass_tflops"], "peak": peak, "unit": "TFLOP/s", "frac": o["roofline_frac"],
                "peak_source": "cuBLAS FP64 GEMM 4096^3 measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry; "
                               f"its HBM figure is {hbm} GB/s); DMMA issue ceiling measured 37.1 TFLOP/s (profiles/)",
                "algorithmic_flops_per_launch": o["flops"]["subjects"], "kernel_ms": subj_ms,
                "kernel_share_of_step": subj_ms / o["ms_per_step"], "traffic": traffic,
                "traffic_unit": "bytes per launch (ncu, profiles/)", "phase_ms": o["phase_ms"],
                "step_algorithmic_flops": o["flops"]["step"], "step_tflops": o["step_tflops"]}
        line = {"metric": METRIC, "value": o["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": o["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(cfg, spb, world, b),
                "impl_details": {"subjects_per_gpu": spb,
                                 "sharding": (f"subjects across ranks, SVGP statistics summed by {'the peer-memory kernel over NVLink (symmetric memory)' if args.exchange == 'p2p' else 'NCCL all-reduce'}"
                                              if world > 1 else "single GPU"),
                                 "tail": o["tail"], "exchange_note": exchange_note,
                                 "l2": "flushed between timed steps (256 MiB write, outside the per-step event pair)",
                                 "kernel_path": o["kernel_path"]},
                "e2e": o["e2e"], "gpu_launches": o["gpu_launches"], "roofline": roof, "cpu_baseline": cpu,
                "gpu_reference": gpu_ref, "clocks": o["clocks"], "finite": o["finite"], "parity": o["parity"]}
        if lat:
            line["latency_point"] = lat
        if others:
            line["other_configs"] = others
        emit(line)
        bad = [n for n, v in [(cfg, o)] + list(others.items()) if isinstance(v.get("parity"), dict) and not v["parity"]["ok"]]
        if bad:
            print(f"bench: PARITY FAILED on {bad}", file=sys.stderr)
            rc = 1
    if dist is not None:
        flag = torch.tensor([rc], dtype=torch.int32, device=device)
        dist.broadcast(flag, 0)
        rc = int(flag.item())
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
