#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-others --no-parity --no-cpu-baseline --no-latency-point"
timeout 300 $CMD > gpurun_out/r02_b_ncu_plain.json 2> gpurun_out/r02_b_ncu_plain.err; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_subjects_fused3|k_prep3" -s 8 -c 2 -o gpurun_out/r02_fused3_v2 -f $CMD > gpurun_out/r02_ncu2.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r02_fused3_v2.ncu-rep > gpurun_out/r02_fused3_v2_ncu_summary.txt 2>&1; cat gpurun_out/r02_fused3_v2_ncu_summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-others > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_h.err
