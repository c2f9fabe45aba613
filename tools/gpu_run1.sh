#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-others --no-parity --no-cpu-baseline --no-latency-point > gpurun_out/r02_bench_l.json 2> gpurun_out/r02_bench_l.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_l.json')); print(round(d['value']), round(d['ms_per_step'],3), {k: round(v,4) for k,v in d['roofline']['phase_ms'].items()})"
