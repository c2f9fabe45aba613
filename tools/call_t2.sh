set -x
LVAE_PREP3_LONG=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cfg4 or ragged or fresh" 2>&1 | tail -3
LVAE_PREP3_LONG=1 python bench.py --cfg cfg4 --spb 2000 --steps 5 --warmup 3 --no-cpu-baseline --no-latency-point > gpurun_out/bench_cfg4_t2.json 2> gpurun_out/bench_cfg4_t2.err
python bench.py --cfg cfg4 --spb 2000 --steps 5 --warmup 3 --no-cpu-baseline --no-latency-point > gpurun_out/bench_cfg4_t.json 2> gpurun_out/bench_cfg4_t.err
python - <<'PY'
import json
for f in ("bench_cfg4_t","bench_cfg4_t2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["roofline"]["phase_ms"])
    except Exception as e: print(f, "ERR", e)
PY
