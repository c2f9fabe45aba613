set -x
for n in 8 4; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2981$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_y_n${n}_cfg2.json 2> gpurun_out/bench_y_n${n}_cfg2.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29833 bench.py --gpus 8 --cfg cfg3 --spb 1000 --steps 3 --warmup 3 > gpurun_out/bench_y_n8_cfg3.json 2> gpurun_out/bench_y_n8_cfg3.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29834 bench.py --gpus 8 --cfg cfg4 --spb 2500 --steps 5 --warmup 3 > gpurun_out/bench_y_n8_cfg4.json 2> gpurun_out/bench_y_n8_cfg4.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29835 bench.py --gpus 8 --cfg cfg5 --spb 2000 --steps 3 --warmup 3 > gpurun_out/bench_y_n8_cfg5.json 2> gpurun_out/bench_y_n8_cfg5.err
python - <<'PY'
import json
for f in ("bench_y_n8_cfg2","bench_y_n4_cfg2","bench_y_n8_cfg3","bench_y_n8_cfg4","bench_y_n8_cfg5"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["n_gpus"], round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["config"].get("exchange_note"))
    except Exception as e: print(f, "ERR", e)
PY
