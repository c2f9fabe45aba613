set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --cfg cfg4 --spb 2000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_w.json 2> gpurun_out/bench_cfg4_w.err
python - <<'PY'
import json
for f in ("bench_cfg4_w",):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["config"]["kernel_path"], d["roofline"]["phase_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/bench_cfg4_w.err
