set -x
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n${n}_cfg2.json 2> gpurun_out/bench_n${n}_cfg2.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --cfg cfg5 --spb 2000 --steps 3 --warmup 3 > gpurun_out/bench_n8_cfg5.json 2> gpurun_out/bench_n8_cfg5.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --cfg cfg4 --spb 2500 --steps 3 --warmup 3 > gpurun_out/bench_n8_cfg4.json 2> gpurun_out/bench_n8_cfg4.err
python - <<'PY'
import json
for f in ("bench_n8_cfg2","bench_n4_cfg2","bench_n8_cfg5","bench_n8_cfg4"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/bench_n8_cfg2.err
