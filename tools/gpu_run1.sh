#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/r02_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-others --no-cpu-baseline --no-latency-point > gpurun_out/r02_bench_l.json 2> gpurun_out/r02_bench_l.err; echo "rc=$?"; tail -2 gpurun_out/r02_bench_l.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_l.json')); print(round(d['value']), round(d['ms_per_step'],3), {k: round(v,4) for k,v in d['roofline']['phase_ms'].items()}, d['parity']['ok'], d['parity']['vs_exact_rank0']['ours'])"
