#!/bin/bash
# one full ncu capture of the hot kernel on c3 (launch 6 = the small-stage instance of the 4th step)
mkdir -p gpurun_out
FPX_DEBUG_ABLATE=${DBG:-0} timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${KERNEL:-search_find_kernel}" -s ${SKIP:-6} -c 1 -f -o gpurun_out/${OUT:-prof_find} python bench.py --workload c3 --steps 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_find.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_find.log | cut -c1-300
