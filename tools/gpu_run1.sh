#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py::test_full_size_parity_vs_oracle_on_device > gpurun_out/r02_pytest_gpu.log 2>&1; tail -8 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err; echo "bench rc=$?"; tail -5 gpurun_out/r02_bench_g.err
