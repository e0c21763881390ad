#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for dbg in 0 32; do
  FPX_DEBUG_ABLATE=$dbg timeout 300 python bench.py --workload c3 --steps 5 --no-cpu-baseline > gpurun_out/var_$dbg.json 2> gpurun_out/var_$dbg.log
  echo "variant=$dbg"; python tools/show_bench.py gpurun_out/var_$dbg.json
done
