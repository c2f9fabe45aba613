#!/bin/bash
# Large minibatches per GPU (memory layout / index-width sanity at BASELINE's full sizes): gpurun -- 'bash tools/big_batches.sh'
set -x
for c in "cfg2 20000" "cfg4 10000" "cfg5 20000" "cfg3 4000"; do set -- $c
timeout 300 python bench.py --cfg $1 --spb $2 --steps 2 --warmup 3 --no-cpu-baseline --no-latency-point > gpurun_out/bench_big_$1.json 2> gpurun_out/bench_big_$1.err || tail -5 gpurun_out/bench_big_$1.err
done
python - <<'PY'
import json
for f in ("cfg2", "cfg4", "cfg5", "cfg3"):
    try:
        d = json.load(open(f"gpurun_out/bench_big_{f}.json"))
        print(f, d["config"]["subjects_per_gpu"], round(d["value"]), "subjects/s", round(d["ms_per_step"], 2), "ms/step e2e", round(d["e2e"]["value"]), "finite", d["finite"])
    except Exception as e:
        print(f, "ERR", e)
PY
nvidia-smi --query-gpu=memory.used --format=csv
