#!/bin/bash
# round 2, session M: where the Zipf workload's time goes (source-level ncu of the exact kernel), phase timers of the hot kernel
mkdir -p gpurun_out
timeout 300 python tools/sweep.py --workload c3 --steps 5 --variants 0,512 > gpurun_out/sweep_c3_timers.log 2>&1; grep -E "variant|fpx dbg|rror" gpurun_out/sweep_c3_timers.log | tail -6
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"search_smem_kernel<.*15" -s 2 -c 1 -f -o gpurun_out/prof_c5 python bench.py --workload c5 --steps 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c5.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_c5.log | cut -c1-300
ls -la gpurun_out/prof_c5.ncu-rep
