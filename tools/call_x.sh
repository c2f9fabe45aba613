set -x
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_x_n2.json 2> gpurun_out/bench_x_n2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --cfg cfg4 --spb 2000 --steps 5 --warmup 3 > gpurun_out/bench_x_n2_cfg4.json 2> gpurun_out/bench_x_n2_cfg4.err
python - <<'PY'
import json
for f in ("bench_x_n2","bench_x_n2_cfg4"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["config"]["sharding"][:60], d["config"].get("exchange_note"))
    except Exception as e: print(f, "ERR", e)
PY
tail -4 gpurun_out/bench_x_n2_cfg4.err
