set -x
ncu --set full --clock-control none --import-source on -k regex:"k_prep_warp|k_subjects_fused$|k_subjects_fused<" -s 2 -c 2 -o gpurun_out/prof_cfg4_long -f python bench.py --cfg cfg4 --spb 2000 --steps 1 --warmup 1 --no-cpu-baseline --no-latency-point > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof_cfg4_long.ncu-rep; tail -3 gpurun_out/ncu_full.log
