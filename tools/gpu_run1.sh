#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-others --no-cpu-baseline --no-latency-point > gpurun_out/r02_bench_j.json 2> gpurun_out/r02_bench_j.err; echo "cfg2 rc=$?"; tail -2 gpurun_out/r02_bench_j.err
timeout 600 python bench.py --cfg cfg4 --steps 10 --warmup 3 --no-others --no-cpu-baseline --no-latency-point --no-parity > gpurun_out/r02_b_cfg4_split.json 2> gpurun_out/r02_b_cfg4_split.err; echo "cfg4 split rc=$?"
timeout 600 python bench.py --cfg cfg4 --steps 10 --warmup 3 --no-others --no-cpu-baseline --no-latency-point --no-parity --no-split > gpurun_out/r02_b_cfg4_nosplit.json 2> gpurun_out/r02_b_cfg4_nosplit.err; echo "cfg4 nosplit rc=$?"; tail -2 gpurun_out/r02_b_cfg4_nosplit.err
python - <<'PY'
import json
for f in ("r02_bench_j", "r02_b_cfg4_split", "r02_b_cfg4_nosplit"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "plain", round(d["e2e"]["plain_loop_value"]), {k: round(v, 3) for k, v in d["roofline"]["phase_ms"].items()})
PY
