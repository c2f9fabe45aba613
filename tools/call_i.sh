set -x
python bench.py --no-cpu-baseline > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err
cat gpurun_out/bench_i.json | cut -c1-1400; tail -5 gpurun_out/bench_i.err
