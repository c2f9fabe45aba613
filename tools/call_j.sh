python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1
tail -60 gpurun_out/e2e_breakdown.log
