#!/usr/bin/env python
"""bench.py — GP-prior ELBO-step throughput (subjects/sec) of the L-VAE hot path on B200.

    python bench.py --gpus N --steps K --warmup W              (N > 1: launched by torch.distributed.run, one rank/GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W     (the CPU reference arm, rank 0 only)

A "step" is one Hensman minibatch of the path: kernel construction from covariates, batched Cholesky/inverse of the
per-subject blocks and of K_mm/H, the minibatch KL upper bound, ALL its gradients (mu, log_v, kernel hyper-parameters,
noise) and the natural-gradient update of (m, H)  (elbo_functions.py:144-216 + backward + training.py:129-135).
Workload at N=1 = BASELINE.json configs[1]: Health-MNIST-shaped synthetic data, 1000 subjects x 20 time points, L=32,
M=60, full additive kernel (cat(id) + SE(age) + id x age + gender x age + disease x disease_time), all 1000 subjects in
one minibatch per step.  N>1: weak scaling — every rank holds its own 1000 subjects of a N*1000-subject minibatch and
the SVGP sufficient statistics are all-reduced over NCCL between the subject pass and the tail.

`value` is device-resident throughput (inputs in HBM, CUDA events, max over ranks); `e2e` is the same step through the
public Python API (lvae_b200.elbo_functions.minibatch_KLD_upper_bound + backward + natural_gradient_step) with pinned
HOST inputs copied in and the loss and encoder gradients copied out inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gp_prior_elbo_step_subjects_per_sec"
UNIT = "subjects/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cfg", default="cfg2")
    ap.add_argument("--spb", type=int, default=1000, help="subjects per minibatch per GPU")
    ap.add_argument("--L", type=int, default=None)
    ap.add_argument("--M", type=int, default=None)
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 generic kernels, 2 fused DMMA kernel")
    ap.add_argument("--cpu-subjects", type=int, default=100, help="subjects in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["nccl", "p2p"],
                    help="N > 1: statistics exchange by NCCL all-reduce or by the peer-memory kernel (symmetric memory)")
    ap.add_argument("--no-latency-point", action="store_true", help="skip the spb=20 point (the reference's default batch)")
    return ap.parse_args()


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
# CPU arm: the oracle port of the reference, all host threads
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(b, n_subjects, device="cpu"):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lvae_oracle as orc
    L, M = b.L, b.M
    rows = int(b.offsets[n_subjects])
    x, mu0, lv0 = b.x[:rows], b.mu[:rows], b.log_v[:rows]
    k0, k1 = orc.parse_kernel_lists(L, **b.lists, id_covariate=2)
    params = []
    for c in k0 + k1:
        c.outputscale = c.outputscale.to(device).clone().requires_grad_(True)
        params.append(c.outputscale)
        for k in list(c.lengthscales):
            c.lengthscales[k] = c.lengthscales[k].to(device).clone().requires_grad_(True)
            params.append(c.lengthscales[k])
    noise = torch.ones(L, dtype=torch.float64, device=device)
    ragged = isinstance(b.T, tuple)
    state = {"m": b.m.clone(), "H": b.H.clone()}

    def step():
        mu = mu0.clone().requires_grad_(True)
        lv = lv0.clone().requires_grad_(True)
        if ragged:
            kld, gm, gH = orc.kld_iter(k0, k1, noise, L, state["m"], state["H"], x, mu, lv, b.z, b.P, n_subjects, b.N,
                                       True, 2, 1e-6)
        else:
            kld, gm, gH = orc.kld_fixed_T(k0, k1, noise, L, state["m"], state["H"], x, mu, lv, b.z, b.P, n_subjects, b.T,
                                          True, 1e-6)
        kld.sum().backward()
        m1, H1 = orc.ng_step(state["m"], state["H"], gm.detach(), gH.detach(), 1e-3)
        for p_ in params:
            p_.grad = None
        return float(kld.detach().sum())
    return step


def time_gpu_reference(b, device, steps=3, warmup=2):
    """The same oracle port of the reference, run with stock PyTorch CUDA ops (ATen / cuBLAS / cuSOLVER) on this GPU: the
    "existing GPU path" (SURVEY 8d).  All subjects of the minibatch, fwd + backward + NG update."""
    import copy
    bg = copy.copy(b)
    for k in ("x", "mu", "log_v", "z", "m", "H"):
        setattr(bg, k, getattr(b, k).to(device))
    with torch.device(device):
        step = cpu_reference_step_fn(bg, b.P, device=device)
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


def time_cpu(b, n_subjects, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_reference_step_fn(b, n_subjects)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n_subjects / dt, dt


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


def main():
    args = parse()
    _quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = args.cfg

    if args.impl == "reference":
        if rank != 0:
            return 0
        from lvae_b200 import synth
        b = synth.make_batch(cfg, P=args.spb, L=args.L, M=args.M)
        n_sub = min(args.cpu_subjects, args.spb)
        steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
        val, dt = time_cpu(b, n_sub, steps, warm)
        cores = os.cpu_count() or 1
        sample = (f"{n_sub} of the {args.spb} subjects of one minibatch per step (fwd + backward + NG update), "
                  f"{steps} steps after {warm} warm-up, torch CPU FP64 with {cores} threads")
        line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": {"workload": f"{cfg}: Health-MNIST-shaped synthetic, {args.spb} subjects x T={b.T}, L={b.L}, "
                                       f"M={b.M}, additive kernel {len(b.lists['cat_int_kernel']) + len(b.lists['cat_kernel']) + len(b.lists['sqexp_kernel']) + len(b.lists['bin_kernel'])} components"},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    from lvae_b200 import _lib, ops, synth
    import lvae_b200.elbo_functions as EF
    from lvae_b200.training import natural_gradient_step
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200.likelihoods import GaussianLikelihood
    from lvae_b200.constraints import GreaterThan
    from lvae_b200.spec import build_structure, flatten
    lib = _lib.load()
    EF.set_kernel_path(args.path)

    b = make_problem(cfg, args.spb, rank, world, args.L, args.M)
    L, M, Q, P_b, N_b = b.L, b.M, b.x.shape[1], args.spb, b.N
    P_tot = P_b * world
    Tl = np.diff(b.offsets)
    ragged = isinstance(b.T, tuple)
    cm0, cm1 = generate_kernel_batched(L, **b.lists, id_covariate=2)
    cm0, cm1 = cm0.double().to(device), cm1.double().to(device)
    lik = GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=GreaterThan(1e-8)).double().to(device)
    lik.noise = 1.0
    st, ls, os_ = build_structure(flatten(cm0), flatten(cm1), L, device=device)
    ls, os_ = ls.detach(), os_.detach()
    noise = torch.ones(L, dtype=torch.float64, device=device)
    dev = lambda t: t.to(device)
    x, mu, lv, z = dev(b.x), dev(b.mu), dev(b.log_v), dev(b.z)
    m, H = dev(b.m).clone(), dev(b.H).clone()
    offsets = torch.from_numpy(b.offsets).to(torch.int32).to(device)
    T_max, sum_T2 = int(Tl.max()), int((Tl * Tl).sum())
    const = L * (N_b * world) / 2 if ragged else L * P_tot * int(b.T) / 2
    call = (ops.make_kld_call(st, L, M, Q, Tl, device, natural_gradient=True, path=args.path) if ragged else
            ops.KldCall(st, L, M, Q, P_b, N_b, T_max, sum_T2, device, natural_gradient=True, path=args.path))
    is_split = isinstance(call, ops.SplitKldCall)
    ng_ws = torch.empty(int(lib.lvae_ng_workspace_doubles(L, M)), dtype=torch.float64, device=device)
    ng_info = torch.zeros(4, dtype=torch.int32, device=device)
    lr = 1e-3

    exchange_note = None
    if dist is not None and args.exchange == "p2p":       # all ranks agree on the exchange path before any timed work
        ok = 1
        try:
            from lvae_b200 import distributed as D
            D.peer_stats(dist.group.WORLD, call.stats.numel(), device)
        except Exception as ex:
            ok, exchange_note = 0, f"symmetric memory unavailable ({repr(ex)[:120]}): NCCL all-reduce"
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            args.exchange = "nccl"
            exchange_note = exchange_note or "a peer could not map symmetric memory: NCCL all-reduce"

    def device_step():
        call.bind(x, offsets, mu, lv, z, m.view(L, M), H, ls, os_, noise, P_tot / (P_b * world), const, 1e-6)
        call.head()
        EF.exchange_stats(call, dist.group.WORLD if dist is not None else None, args.exchange)
        call.tail()
        rc = lib.lvae_ng_step_f64(_lib.ptr(m), _lib.ptr(H), _lib.ptr(call.grad_m), _lib.ptr(call.grad_H),
                                  _lib.ptr(call.Hinv), lr, L, M, _lib.ptr(ng_ws), _lib.ptr(ng_info),
                                  _lib.stream_ptr(device))
        _lib.check(rc, "lvae_ng_step_f64")

    if args.path == 1:
        kernel_path = "generic"
    elif M <= 64:
        kernel_path = ("fused (one DMMA kernel per subject pass)" if T_max <= 24 or args.path == 2 else
                       "split: subjects with T <= 24 fused v2 + prep v3, longer ones fused v1 + 4-warp prep" if
                       is_split else "fused v1")
    else:
        kernel_path = "gemm (U/V materialised, S = U^T U and Y = V W as batched DMMA GEMMs)" if T_max <= 24 else "generic"
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=device)   # > 126 MB L2
    fl = algorithmic_flops(Tl, L, M, st.n_comp0, st.n_comp1, P_b)
    peak = dgemm_peak_tflops(device) if rank == 0 else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident timing -------------------------------------------------------------------------------
    m0, H0 = m.clone(), H.clone()
    for _ in range(args.warmup):
        device_step()
    barrier()
    lib.lvae_profile_enable(1)
    launches0 = ops.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    phase_ms = np.zeros((args.steps, 6))
    clocks = ClockSampler(local_rank)
    if rank == 0:                                  # rank 0's GPU, sampled through the device-timed AND the end-to-end timed regions
        clocks.__enter__()
    barrier()
    for i in range(args.steps):
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
    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = P_b * world / (ms_per_step * 1e-3)
    finite = bool(torch.isfinite(call.kld_per_latent).all() and torch.isfinite(H).all())

    # ---- end to end through the public API with host buffers --------------------------------------------------
    del call                                   # its workspace (tens of GB at M > 64 with large minibatches) is not needed any more
    torch.cuda.empty_cache()
    m.copy_(m0); H.copy_(H0)
    hx, hmu, hlv = b.x.pin_memory(), b.mu.pin_memory(), b.log_v.pin_memory()
    out_mu = torch.empty_like(b.mu).pin_memory()
    out_lv = torch.empty_like(b.log_v).pin_memory()
    out_kld = torch.empty(1, dtype=torch.float64).pin_memory()
    if dist is not None:
        EF.set_process_group(dist.group.WORLD, args.exchange)
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

    def e2e_step(i):
        d = dbuf[i % 2]
        prefetch(i + 1)
        main.wait_event(d["ready"])
        xd = d["x"]
        mud = d["mu"].detach().requires_grad_(True)
        lvd = d["lv"].detach().requires_grad_(True)
        if ragged:
            kld, gm, gH = EF.minibatch_KLD_upper_bound_iter(cm0, cm1, lik, L, state["m"], state["H"], xd, mud, lvd, z,
                                                            P_tot, P_b * world, N_b * world, True, 2, 1e-6)
        else:
            kld, gm, gH = EF.minibatch_KLD_upper_bound(cm0, cm1, lik, L, state["m"], state["H"], xd, mud, lvd, z, P_tot,
                                                       P_b * world, int(b.T), True, 1e-6)
        kld.sum().backward()
        d["free"].record(main)
        state["m"], state["H"] = natural_gradient_step(state["m"], state["H"], gm, gH, lr)
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
        cm0.zero_grad(set_to_none=True); cm1.zero_grad(set_to_none=True)

    for d in dbuf:
        d["free"].record(main)
    n_e2e = 0
    prefetch(0)
    for _ in range(max(3, args.warmup)):
        e2e_step(n_e2e); n_e2e += 1
    main.wait_event(out_done)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step(n_e2e); n_e2e += 1
    main.wait_event(out_done)                              # the last step's results have reached host memory
    e1.record()
    barrier()
    if rank == 0:
        clocks.__exit__(None, None, None)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / args.steps
    e2e_val = P_b * world / (e2e_ms * 1e-3)
    h2d = (hx.numel() + hmu.numel() + hlv.numel()) * 8
    d2h = (out_mu.numel() + out_lv.numel() + 1) * 8
    EF.check_errors()
    EF.set_error_check("immediate")
    EF.set_process_group(None)

    lat = None
    if not args.no_latency_point and rank == 0 and world == 1 and P_b >= 20:
        spb2 = 20
        rows = int(b.offsets[spb2])
        off2 = offsets[:spb2 + 1].contiguous()
        c2 = ops.KldCall(st, L, M, Q, spb2, rows, T_max, int((Tl[:spb2] ** 2).sum()), device, True, args.path)
        mm, HH = m0.clone(), H0.clone()

        def small():
            c2.bind(x[:rows], off2, mu[:rows], lv[:rows], z, mm.view(L, M), HH, ls, os_, noise, P_tot / spb2, const, 1e-6)
            c2.run()
            lib.lvae_ng_step_f64(_lib.ptr(mm), _lib.ptr(HH), _lib.ptr(c2.grad_m), _lib.ptr(c2.grad_H),
                                 _lib.ptr(c2.Hinv), lr, L, M, _lib.ptr(ng_ws), _lib.ptr(ng_info),
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
        try:      # the same small minibatch through the public API with pinned host inputs (Python + launch overhead regime)
            hx2, hmu2, hlv2 = b.x[:rows].pin_memory(), b.mu[:rows].pin_memory(), b.log_v[:rows].pin_memory()
            st2 = {"m": m0.clone(), "H": H0.clone()}
            o_mu = torch.empty_like(b.mu[:rows]).pin_memory()

            def small_api():
                xd = hx2.to(device, non_blocking=True)
                mud = hmu2.to(device, non_blocking=True).requires_grad_(True)
                lvd = hlv2.to(device, non_blocking=True).requires_grad_(True)
                if ragged:
                    kld, gm, gH = EF.minibatch_KLD_upper_bound_iter(cm0, cm1, lik, L, st2["m"], st2["H"], xd, mud, lvd, z, P_tot,
                                                                    spb2, N_b, True, 2, 1e-6)
                else:
                    kld, gm, gH = EF.minibatch_KLD_upper_bound(cm0, cm1, lik, L, st2["m"], st2["H"], xd, mud, lvd, z, P_tot, spb2,
                                                               int(b.T), True, 1e-6)
                kld.sum().backward()
                st2["m"], st2["H"] = natural_gradient_step(st2["m"], st2["H"], gm, gH, lr)
                o_mu.copy_(mud.grad, non_blocking=True)
                cm0.zero_grad(set_to_none=True); cm1.zero_grad(set_to_none=True)
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
                mg, Hg = m0.clone().reshape(L, M, 1).contiguous(), H0.clone().contiguous()
                gstep = GraphedHensmanStep(cm0, cm1, lik, L, mg, Hg, z, P_tot, spb2, int(b.T), 1e-6, lr)

                def small_graphed():
                    xd = hx2.to(device, non_blocking=True)
                    mud = hmu2.to(device, non_blocking=True).requires_grad_(True)
                    lvd = hlv2.to(device, non_blocking=True).requires_grad_(True)
                    gstep(xd, mud, lvd).backward()
                    o_mu.copy_(mud.grad, non_blocking=True)
                    cm0.zero_grad(set_to_none=True); cm1.zero_grad(set_to_none=True)
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

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            n_sub = min(args.cpu_subjects, args.spb)
            v, dt = time_cpu(make_problem(cfg, args.spb, 0, 1, args.L, args.M), n_sub, 3, 1)
            cores = os.cpu_count() or 1
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"oracle port of the reference (torch CPU FP64, {cores} threads): {n_sub} of the {args.spb} "
                             f"subjects per step, fwd + backward + NG update, 3 steps after 1 warm-up ({dt:.2f} s/step)"}
        gpu_ref = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, ms = time_gpu_reference(make_problem(cfg, args.spb, 0, 1, args.L, args.M), device)
                gpu_ref = {"value": v, "unit": UNIT, "ms_per_step": ms,
                           "what": "oracle port of the reference with stock PyTorch CUDA ops (ATen/cuBLAS/cuSOLVER) on this "
                                   "GPU, device-resident inputs, all subjects of the minibatch, fwd + backward + NG update"}
            except Exception as ex:   # the baseline is informative only
                gpu_ref = {"unavailable": repr(ex)[:200]}
        subj_ms = float(phase_ms[:, 2].mean())
        achieved = fl["subjects"] / (subj_ms * 1e-3) * 1e-12
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else None
        roof_kernel = ("subject pass (fused kernel blocks + trisolve + S/Y contractions, FP64 DMMA)" if M <= 64 else
                       "subject pass (k_uv + S = U^T U GEMM + Y = V W GEMM + k_adj, FP64 DMMA)")
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tfile) and M <= 64 and T_max <= 24:
            ent = json.load(open(tfile)).get(f"{cfg}:spb{P_b}:k_subjects_fused2")
            traffic = ent["traffic_bytes"] if ent else None
        roof = {"bound": "tensor", "kernel": roof_kernel,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": "cuBLAS FP64 GEMM 4096^3 measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry; "
                               f"its HBM figure is {hbm} GB/s); DMMA issue ceiling measured 37.1 TFLOP/s (profiles/)",
                "algorithmic_flops_per_launch": fl["subjects"], "kernel_ms": subj_ms,
                "kernel_share_of_step": subj_ms / ms_per_step, "traffic": traffic, "traffic_unit": "bytes per launch (ncu, profiles/)",
                "phase_ms": {n: float(phase_ms[:, i].mean()) for i, n in
                             enumerate(["head", "prep", "subjects", "reduce", "tail", "ng_step"])},
                "step_algorithmic_flops": fl["step"], "step_tflops": fl["step"] / (ms_per_step * 1e-3) * 1e-12}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{cfg}: Health-MNIST-shaped synthetic, {P_b} subjects/GPU x T={b.T}, L={L}, M={M}, "
                                       f"{st.n_comp0}+{st.n_comp1} additive components, one minibatch of {P_b * world} subjects "
                                       f"per step (bound + all gradients + NG update)",
                           "subjects_per_gpu": P_b, "global_batch_subjects": P_b * world, "L": L, "M": M,
                           "sharding": (f"subjects across ranks, SVGP statistics summed by {'the peer-memory kernel over NVLink (symmetric memory)' if args.exchange == 'p2p' else 'NCCL all-reduce'}"
                                        if world > 1 else "single GPU"),
                           "exchange_note": exchange_note,
                           "l2": "flushed between timed steps (256 MiB write, outside the per-step event pair)",
                           "kernel_path": kernel_path},
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms,
                        "how": "public API (minibatch_KLD_upper_bound + backward + natural_gradient_step); every step copies "
                               "x, mu, log_v from pinned host memory and returns kld, d_mu, d_log_v to pinned host memory; the "
                               "copies of neighbouring steps overlap the compute on two copy streams"},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "gpu_reference": gpu_ref,
                "clocks": clocks.summary(),
                "finite": finite}
        if lat:
            line["latency_point"] = lat
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
