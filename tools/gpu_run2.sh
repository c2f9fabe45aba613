#!/bin/bash
# 2-GPU pass: multi-GPU tests and the bench at N = 2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02_pytest_multi.log 2>&1; tail -8 gpurun_out/r02_pytest_multi.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r02_bench_n2.err
