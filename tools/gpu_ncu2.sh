#!/bin/bash
mkdir -p gpurun_out
FPX_DEBUG_ABLATE=${DBG:-1024} timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KERNEL:-search_sketch2_kernel}" -s 3 -c 1 -f -o gpurun_out/${OUT:-prof_s2} python bench.py --workload c3 --steps 1 --no-cpu-baseline > gpurun_out/ncu_s2.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_s2.log | cut -c1-300
