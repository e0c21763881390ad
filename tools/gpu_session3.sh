#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench c3"; timeout 900 python bench.py --workload c3 --steps 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; echo "rc=$?"; tail -4 gpurun_out/bench_c3.log; cat gpurun_out/bench_c3.json
