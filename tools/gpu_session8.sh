#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"search_sketch_kernel" -s 3 -c 1 -o gpurun_out/prof_v13 python bench.py --workload c3 --steps 1 --no-cpu-baseline > gpurun_out/ncu_v13.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 40 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 4 --no-cpu-baseline > gpurun_out/launches.log 2>&1; echo "launch list rc=$?"
