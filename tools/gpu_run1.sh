#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_final_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_final_ref.json 2> gpurun_out/r02_bench_final_ref.err; echo "ref rc=$?"
