#!/bin/bash
# round 2, session J: GPU tests, then the other single-GPU workloads (c5 Zipf, c2)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wl in c5 c2; do
  timeout 400 python bench.py --workload $wl --steps 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.log; echo "$wl rc=$?"; python tools/show_bench.py gpurun_out/bench_$wl.json; grep -iE "error|Traceback" -A8 gpurun_out/bench_$wl.log | head -20
done
