set -x
ncu --set full --clock-control none --import-source on -k regex:"k_prep3" -s 6 -c 1 -o gpurun_out/prof_prep3 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency-point > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_prep3" -s 3 -c 1 -o gpurun_out/prof_prep3_cfg4 -f python bench.py --cfg cfg4 --spb 2000 --steps 1 --warmup 1 --no-cpu-baseline --no-latency-point > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/prof_prep3*.ncu-rep
