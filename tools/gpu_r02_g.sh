#!/bin/bash
# round 2, session G: quick sweep of the hot kernel (no tests)
mkdir -p gpurun_out
timeout 200 python tools/sweep.py --workload c3 --steps 5 --variants ${VARIANTS:-0,0x1000000,0x2000000,512} --check ${CHECK:-0x1000000,0x2000000} --out gpurun_out/sweep_c3.txt > gpurun_out/sweep_c3.log 2>&1
echo "sweep c3 rc=$?"; grep -E "variant|fpx dbg|rror" gpurun_out/sweep_c3.log | tail -40
