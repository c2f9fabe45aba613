for rep in 1 2; do for un in 0 1; do
LVAE_PREP3_TU=1 LVAE_PREP3_UNR=$un python bench.py --no-cpu-baseline --no-latency-point --steps 50 > gpurun_out/bench_un${un}.json 2> gpurun_out/bench_tu.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_un${un}.json")); print("rep", $rep, "entry-unroll", $un, round(d["value"]), round(d["roofline"]["phase_ms"]["prep"],4), round(d["roofline"]["phase_ms"]["subjects"],4))
PY
done; done
