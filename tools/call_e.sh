set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py --cfg cfg3 --spb 1000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_c.json 2> gpurun_out/bench_cfg3_c.err
python bench.py --cfg cfg5 --spb 2000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_c.json 2> gpurun_out/bench_cfg5_c.err
for c in "cfg3 1000" "cfg5 2000"; do set -- $c
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_uv|k_adj|k_gemm|k_prep_warp" -c 60 --csv --log-file gpurun_out/launches_$1.csv python bench.py --cfg $1 --spb $2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1
done
