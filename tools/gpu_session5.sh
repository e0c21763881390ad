#!/bin/bash
mkdir -p gpurun_out
python tools/debug_sketch.py > gpurun_out/debug_sketch.log 2>&1; cat gpurun_out/debug_sketch.log | tail -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_sketch_kernel -s 3 -c 1 -o gpurun_out/prof_sketch_v3 python bench.py --workload c3 --steps 1 --no-cpu-baseline > gpurun_out/ncu_v3.log 2>&1; echo "ncu rc=$?"
