set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_a.json 2> gpurun_out/bench_ref_a.err
python bench.py --cfg cfg3 --spb 1000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_a.json 2> gpurun_out/bench_cfg3_a.err
python bench.py --cfg cfg5 --spb 2000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_a.json 2> gpurun_out/bench_cfg5_a.err
python bench.py --cfg cfg4 --spb 2000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_a.json 2> gpurun_out/bench_cfg4_a.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
tail -3 gpurun_out/pytest_gpu_a.log
cat gpurun_out/bench_a.json
