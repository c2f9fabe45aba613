set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --no-cpu-baseline --no-latency-point > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
python bench.py --cfg cfg4 --spb 2000 --steps 5 --warmup 3 --no-cpu-baseline --no-latency-point > gpurun_out/bench_cfg4_t.json 2> gpurun_out/bench_cfg4_t.err
python bench.py --cfg cfg5 --spb 2000 --steps 3 --warmup 3 --no-cpu-baseline --no-latency-point > gpurun_out/bench_cfg5_t.json 2> gpurun_out/bench_cfg5_t.err
python bench.py --cfg cfg3 --spb 1000 --steps 3 --warmup 3 --no-cpu-baseline --no-latency-point > gpurun_out/bench_cfg3_t.json 2> gpurun_out/bench_cfg3_t.err
python - <<'PY'
import json
for f in ("bench_t","bench_cfg4_t","bench_cfg5_t","bench_cfg3_t"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["roofline"]["phase_ms"])
    except Exception as e: print(f, "ERR", e)
PY
