#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench n$N rc=$?"; grep -v "^\[W\|^W1018\|NCCL\|^$" gpurun_out/r02_bench_n$N.err | tail -5
