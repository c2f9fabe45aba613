#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/r02_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu.log
timeout 600 python tools/widening_bench.py > gpurun_out/r02_widening_bench.json 2> gpurun_out/r02_widening.err; tail -3 gpurun_out/r02_widening.err; cat gpurun_out/r02_widening_bench.json | head -c 1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_dubo.csv python tools/dubo_once.py > gpurun_out/r02_dubo_once.log 2>&1; python tools/launch_list_summary.py gpurun_out/r02_launches_dubo.csv > gpurun_out/r02_launches_dubo_summary.txt 2>&1; head -30 gpurun_out/r02_launches_dubo_summary.txt
