set -x
python bench.py > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_m.json")); print(d["value"], d["e2e"]["value"], d["cpu_baseline"], d["gpu_reference"], d["roofline"]["traffic"], d["clocks"])
PY
tail -3 gpurun_out/bench_m.err
