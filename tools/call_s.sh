set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_s.json")); print(d["value"], d["e2e"]["value"], d["latency_point"], d["clocks"])
PY
tail -3 gpurun_out/bench_s.err
