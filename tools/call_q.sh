set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -25
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --exchange p2p > gpurun_out/bench_n2_p2p.json 2> gpurun_out/bench_n2_p2p.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --exchange nccl > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err
python - <<'PY'
import json
for f in ("bench_n2_p2p","bench_n2_nccl"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["sharding"], d["config"].get("exchange_note"))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/bench_n2_p2p.err
