#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/sweep.py --workload ${WL:-c3} --steps 5 --variants ${VARIANTS} --check ${CHECK:-1024} --out gpurun_out/sweep2.txt > gpurun_out/sweep2.log 2>&1
echo "rc=$?"; grep -E "variant|fpx dbg|rror" gpurun_out/sweep2.log | tail -40
