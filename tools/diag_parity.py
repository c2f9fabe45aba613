"""Diagnostic: bench.check_parity over a sweep of sizes / hyper-parameter modes (prints per-tensor errors and magnitudes)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ARGS = argparse.Namespace(path=int(os.environ.get("DIAG_PATH", "0")), exchange="nccl", no_parity=False)
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
for mode in ("default", "unit_noise", "perturbed"):
    os.environ["LVAE_BENCH_HYPERS"] = mode
    for (cfg, spb, L, M) in [("cfg2", 40, 4, 60), ("cfg2", 40, 32, 60), ("cfg2", 1000, 4, 60), ("cfg2", 1000, 32, 60)]:
        try:
            r = bench.run_config(ARGS, cfg, spb, 0, 1, dev, None, None, steps=1, warmup=0, parity_only=True, L=L, M=M)
            p = r.out["parity"]
            print(mode, cfg, spb, L, M, "max_rel %.2e" % p["max_rel"], {k: "%.1e" % v for k, v in p["per_tensor_rank0"].items()}, flush=True)
        except Exception as ex:
            print(mode, cfg, spb, L, M, "EXC", repr(ex)[:200], flush=True)
