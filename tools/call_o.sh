set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --no-cpu-baseline --latency-point > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err
python bench.py --cfg cfg3 --spb 1000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_o.json 2> gpurun_out/bench_cfg3_o.err
python - <<'PY'
import json
for f in ("bench_o","bench_cfg3_o"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["phase_ms"], d.get("latency_point"))
PY
