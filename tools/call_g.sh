set -x
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_cfg2.json 2> gpurun_out/bench_n2_cfg2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --cfg cfg5 --spb 2000 --steps 3 --warmup 3 > gpurun_out/bench_n2_cfg5.json 2> gpurun_out/bench_n2_cfg5.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err
cat gpurun_out/bench_n2_cfg2.json gpurun_out/bench_n2_cfg5.json gpurun_out/bench_n2_ref.json | cut -c1-700
tail -3 gpurun_out/bench_n2_cfg5.err
