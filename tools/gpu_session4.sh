#!/bin/bash
mkdir -p gpurun_out
echo "== gather ubench"; timeout 300 ./tools/ubench_gather > gpurun_out/ubench_gather2.log 2>&1; cat gpurun_out/ubench_gather2.log
echo "== ncu full of sketch kernel (c3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_sketch_kernel -s 3 -c 1 -o gpurun_out/prof_sketch_v2 python bench.py --workload c3 --steps 1 --no-cpu-baseline > gpurun_out/ncu_v2.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_v2.log | cut -c1-300
ls -la gpurun_out/
