#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"search_sketch_kernel" -s 3 -c 1 -o gpurun_out/prof_v14 python bench.py --workload c3 --steps 1 --no-cpu-baseline > gpurun_out/ncu_v14.log 2>&1; echo "ncu rc=$?"
