#!/bin/bash
# Launch list + one full capture of the two dominant kernels of the headline step (run only after bench.py exited 0 without ncu).
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-others --no-parity --no-cpu-baseline --no-latency-point"
timeout 300 $B > gpurun_out/ncu_pre.json 2> gpurun_out/ncu_pre.err || { echo "bench failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_bench_cfg2.csv $B > /dev/null 2> gpurun_out/ncu_list.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_subjects_fused3|k_prep3" -s 6 -c 2 -f -o gpurun_out/r02_final_cfg2 $B > /dev/null 2> gpurun_out/ncu_full.err
ls -la gpurun_out/r02_final_cfg2.ncu-rep gpurun_out/r02_launches_bench_cfg2.csv
