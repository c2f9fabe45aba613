set -x
python bench.py > gpurun_out/bench_final_cfg2.json 2> gpurun_out/bench_final_cfg2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_final_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency-point > gpurun_out/ncu_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_prep3|k_subjects_fused2" -s 6 -c 2 -o gpurun_out/prof_final_cfg2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency-point > gpurun_out/ncu_full.log 2>&1
for c in "cfg3 1000" "cfg5 2000"; do set -- $c
python bench.py --cfg $1 --spb $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final_$1.json 2> gpurun_out/bench_final_$1.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_uv|k_adj|k_gemm|k_prep|k_diag|k_big|k_reduce|k_plan" -c 120 --csv --log-file gpurun_out/launches_final_$1.csv python bench.py --cfg $1 --spb $2 --steps 1 --warmup 1 --no-cpu-baseline --no-latency-point > gpurun_out/ncu_$1.log 2>&1
done
python - <<'PY'
import json
for f in ("bench_final_cfg2","bench_final_ref","bench_final_cfg3","bench_final_cfg5"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d.get("roofline",{}).get("frac"), d.get("latency_point"))
PY
