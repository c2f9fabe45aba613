set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py --cfg cfg4 --spb 2000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_b.json 2> gpurun_out/bench_cfg4_b.err
python bench.py --no-cpu-baseline > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err
python - <<'PY'
import json
for f in ("bench_cfg4_b","bench_k"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["phase_ms"])
PY
