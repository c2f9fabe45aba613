set -x
python -m pytest tests -m gpu -q 2>&1 | tail -30
python bench.py --cfg cfg3 --spb 1000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_b.json 2> gpurun_out/bench_cfg3_b.err
python bench.py --cfg cfg5 --spb 2000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_b.json 2> gpurun_out/bench_cfg5_b.err
cat gpurun_out/bench_cfg3_b.json gpurun_out/bench_cfg5_b.json; tail -5 gpurun_out/bench_cfg3_b.err
