#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"search_sketch_kernel|prepare_kernel" -s 6 -c 2 -o gpurun_out/prof_v4 python bench.py --workload c3 --steps 1 --no-cpu-baseline > gpurun_out/ncu_v4.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
