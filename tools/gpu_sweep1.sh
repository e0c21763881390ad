#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python tools/sweep.py --workload c3 --steps 5 --variants ${VARIANTS:-0,512,513,514,515,520,528,576,640} --out gpurun_out/sweep1.txt > gpurun_out/sweep1.log 2>&1
echo "rc=$?"; grep -E "variant|fpx dbg" gpurun_out/sweep1.log
