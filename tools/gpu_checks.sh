#!/bin/bash
# What the GPU box is asked to run after a change (use with: gpurun --timeout 900 -- 'bash tools/gpu_checks.sh').
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_check_cfg2.json 2> gpurun_out/bench_check_cfg2.err
for c in "cfg3 1000" "cfg4 2000" "cfg5 2000"; do set -- $c
python bench.py --cfg $1 --spb $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_check_$1.json 2> gpurun_out/bench_check_$1.err
done
python - <<'PY'
import json
for f in ("cfg2", "cfg3", "cfg4", "cfg5"):
    d = json.load(open(f"gpurun_out/bench_check_{f}.json"))
    print(f, round(d["value"]), "subjects/s", round(d["ms_per_step"], 3), "ms/step  e2e", round(d["e2e"]["value"]), d["roofline"]["phase_ms"])
PY
