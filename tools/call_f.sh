set -x
ncu --set full --clock-control none --import-source on -k regex:"k_adj" -s 1 -c 1 -o gpurun_out/prof_adj2 -f python bench.py --cfg cfg5 --spb 2000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
