"""Timing of the widened rows on one GPU at cfg2's shape (P subjects x T=20, L=32, M=60):
validation_dubo forward + backward through the CUDA ops, batch_predict, and — as the "existing GPU path" next to it — the
same Python with every C-ABI op swapped for stock PyTorch CUDA ops (tests/ops_emulation.py).  Device time with CUDA events
after warm-up; writes one JSON line (and gpurun_out/widening_bench.json when that directory exists).

    python tools/widening_bench.py [--P 1000] [--steps 10]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    from lvae_b200 import ops, synth, utils as U
    from lvae_b200.constraints import GreaterThan
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200.likelihoods import GaussianLikelihood
    from lvae_b200.validation import validation_dubo
    from ops_emulation import emulated_ops
    dev = "cuda"
    b = synth.make_batch("cfg2", P=a.P)
    L, T, P = b.L, b.T, b.P
    cm0, cm1 = generate_kernel_batched(L, **b.lists, id_covariate=2)
    cm0, cm1 = cm0.double().to(dev), cm1.double().to(dev)
    lik = GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=GreaterThan(1e-8)).double().to(dev)
    x, z = b.x.to(dev), b.z.to(dev)
    mu, lv = b.mu.to(dev).requires_grad_(True), b.log_v.to(dev).requires_grad_(True)
    params = [mu, lv] + [p for m in (cm0, cm1, lik) for p in m.parameters()]

    def dubo_step():
        for p in params:
            p.grad = None
        validation_dubo(L, cm0, cm1, lik, x, mu, lv, z, P, T, 1e-6).sum().backward()

    def dubo_fwd():
        with torch.no_grad():
            validation_dubo(L, cm0, cm1, lik, x, mu, lv, z, P, T, 1e-6)

    bt = synth.make_batch("cfg2", P=a.P + 100, seed=99)
    test_x = bt.x[-2000:].to(dev)

    def predict():
        with torch.no_grad():
            U.batch_predict(L, cm0, cm1, lik, x, test_x, mu.detach(), z, P, T, 2, 1e-6)

    out = {"workload": f"cfg2 shape: P={P} T={T} L={L} M={b.M}", "steps": a.steps}
    n0 = ops.launch_count()
    dubo_step()
    out["dubo_fwd_bwd_own_launches"] = int(ops.launch_count() - n0)
    out["dubo_fwd_bwd_ms"] = timed(dubo_step, a.steps)
    out["dubo_fwd_ms"] = timed(dubo_fwd, a.steps)
    out["predict_2000_rows_ms"] = timed(predict, a.steps)
    g_own = [p.grad.clone() for p in params]
    with emulated_ops():
        out["stock_torch_dubo_fwd_bwd_ms"] = timed(dubo_step, max(2, a.steps // 3), warmup=1)
        out["stock_torch_dubo_fwd_ms"] = timed(dubo_fwd, max(2, a.steps // 3), warmup=1)
    err = max(float((g - p.grad).abs().max() / p.grad.abs().max().clamp_min(1e-300)) for g, p in zip(g_own, params))
    out["max_rel_grad_diff_vs_stock_torch"] = err
    out["subjects_per_s_dubo_fwd_bwd"] = P / out["dubo_fwd_bwd_ms"] * 1e3
    # the non-Hensman loops' pattern: per-latent un-batched module lists, one deviance_upper_bound call per latent
    # (training.py:334-343) against ONE deviance_upper_bound_all call over the same lists
    import lvae_b200.elbo_functions as EF
    from lvae_b200.kernel_gen import generate_kernel_approx
    trip = []
    for _ in range(L):
        u0, u1 = generate_kernel_approx(**b.lists, id_covariate=2)
        trip.append((u0.double().to(dev), u1.double().to(dev),
                     GaussianLikelihood(noise_constraint=GreaterThan(1e-8)).double().to(dev)))
    c0, c1, lk = [t[0] for t in trip], [t[1] for t in trip], [t[2] for t in trip]
    zl = [z[i].contiguous() for i in range(L)]
    lparams = [mu, lv] + [p for t in trip for mod in t for p in mod.parameters()]

    def loop_step():
        for p in lparams:
            p.grad = None
        tot = 0.0
        for i in range(L):
            tot = tot + EF.deviance_upper_bound(c0[i], c1[i], lk[i], x, mu[:, i], lv[:, i], zl[i], P, T, 1e-6)
        tot.backward()

    def all_step():
        for p in lparams:
            p.grad = None
        EF.deviance_upper_bound_all(c0, c1, lk, x, mu, lv, zl, P, T, 1e-6).sum().backward()

    out["per_latent_loop_dubo_fwd_bwd_ms"] = timed(loop_step, max(2, a.steps // 3), warmup=1)
    out["dubo_all_fwd_bwd_ms"] = timed(all_step, a.steps)
    line = json.dumps(out)
    print(line)
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        open(os.path.join(ROOT, "gpurun_out", "widening_bench.json"), "w").write(line + "\n")


if __name__ == "__main__":
    main()
