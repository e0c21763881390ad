#!/bin/bash
mkdir -p gpurun_out
echo "== gather ubench"; timeout 300 ./tools/ubench_gather > gpurun_out/ubench_gather.log 2>&1; cat gpurun_out/ubench_gather.log
echo "== ncu full of current kernel (c3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_smem_kernelILi14 -s 3 -c 1 -o gpurun_out/prof_v1_c3 python bench.py --workload c3 --steps 1 --no-cpu-baseline > gpurun_out/ncu_v1.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_v1.log
ls -la gpurun_out/
