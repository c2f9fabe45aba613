set -x
ncu --set full --clock-control none --import-source on -k regex:"k_prep_warp|k_subjects_fused2" -s 6 -c 2 -o gpurun_out/prof_cfg2_v10 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof_cfg2_v10.ncu-rep
