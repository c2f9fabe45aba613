set -x
for ex in p2p nccl p2p nccl; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 296$((RANDOM%80+10)) bench.py --gpus 8 --steps 20 --warmup 5 --exchange $ex --no-cpu-baseline > gpurun_out/bench_n8_$ex.json 2> gpurun_out/bench_n8_$ex.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n8_$ex.json")); print("$ex", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("exchange_note"))
except Exception as e: print("$ex", "ERR", e)
PY
done
tail -3 gpurun_out/bench_n8_p2p.err
