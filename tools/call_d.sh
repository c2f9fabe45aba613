set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "big_path" 2>&1 | tail -5
for c in "cfg3 1000" "cfg5 2000"; do set -- $c
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_uv|k_adj|k_gemm|k_prep_warp|k_reduce_big|k_plan|k_diag|k_big" -c 200 --csv --log-file gpurun_out/launches_$1.csv python bench.py --cfg $1 --spb $2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1
done
