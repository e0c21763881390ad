#!/bin/bash
mkdir -p gpurun_out
FPX_DEBUG_ABLATE=9 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"search_sketch_kernel" -s 3 -c 1 -o gpurun_out/prof_v14_skel python bench.py --workload c3 --steps 1 --no-cpu-baseline > gpurun_out/ncu_v14_skel.log 2>&1; echo "ncu rc=$?"
